"""Turn the logs of tests/cuda/round2_validate.sh into a markdown table (stdout) for profiles/r02_validation.md:
one row per step with its exit code and, for bench steps, tokens/s and ms/step from the JSON line.
Usage: python tests/cuda/summarize_round2.py gpurun_out/round2_single.log gpurun_out/round2_multi.log"""
import json
import re
import sys


def steps(path):
    cur, body = None, []
    for line in open(path, errors="replace"):
        line = line.rstrip("\n")
        if line.startswith("=== "):
            if cur is not None:
                yield cur, body, None
            cur, body = line[4:], []
        elif line.startswith("--- exit ") and cur is not None:
            yield cur, body, int(line.split()[-1])
            cur, body = None, []
        elif cur is not None:
            body.append(line)
    if cur is not None:
        yield cur, body, None


def main(paths):
    print("| step | exit | tokens/s | ms/step | note |")
    print("|---|---|---|---|---|")
    for path in paths:
        for cmd, body, rc in steps(path):
            val = ms = note = ""
            for line in body:
                if line.startswith("{") and '"value"' in line:
                    try:
                        d = json.loads(line)
                        val, ms = f"{d['value']:.0f}", f"{d['ms_per_step']:.1f}"
                        note = d.get("config", {}).get("parallelism", "")
                    except Exception:
                        pass
                m = re.search(r"(\d+) passed|(\d+) failed", line)
                if m and not val:
                    note = line.strip()[-60:]
                if "-> OK" in line or "-> FAIL" in line:
                    note = (note + " " + line.split(":")[0].strip() + ("=OK" if "-> OK" in line else "=FAIL")).strip()
                if line.startswith("gemv M="):
                    note = (note + " " + line.strip()).strip()
            short = re.sub(r"\s+", " ", cmd)
            short = short.replace("python -m torch.distributed.run --nnodes=1 ", "torchrun ").replace("--master-addr 127.0.0.1 ", "")
            print(f"| `{short[:150]}` | {rc if rc is not None else 'cut'} | {val} | {ms} | {note[:120]} |")


if __name__ == "__main__":
    main(sys.argv[1:] or ["gpurun_out/round2_single.log", "gpurun_out/round2_multi.log"])
