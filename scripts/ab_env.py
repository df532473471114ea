"""Time RectifiedFlow.sample (base, T=256, 50 steps) under a list of environment-variable settings, one fresh process each,
and compare every result with the first.  Usage: python scripts/ab_env.py B precision NAME:K=V,K=V NAME2:K=V ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.ab_knobs import CHILD  # noqa: E402


def main():
    import torch
    B, prec = int(sys.argv[1]), sys.argv[2]
    ref = None
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for spec in sys.argv[3:]:
        name, _, kv = spec.partition(":")
        env = dict(os.environ)
        for item in filter(None, kv.split(",")):
            k, _, v = item.partition("=")
            env[k] = v
        path = os.path.join(ROOT, "gpurun_out", f"abenv_{name}.pt")
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, str(B), "50", prec, path], env=env,
                           capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print(json.dumps({"knobs": name, "error": r.stderr[-800:]}), flush=True)
            continue
        rec = json.loads(r.stdout.strip().splitlines()[-1])
        out = torch.load(path)
        os.remove(path)
        if ref is None:
            ref = out
        rec.update(knobs=name, streams=B, precision=prec,
                   rel_vs_first=float((out.double() - ref.double()).norm() / ref.double().norm()))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
