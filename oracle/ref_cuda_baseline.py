"""Timed run of the reference's OWN CUDA path (featExtract -d0) on the GPU box (BENCH INFRASTRUCTURE ONLY).

oracle/_ref/featExtract_ref_cuda = the reference's featExtract.cpp + src_common + cuda_common/SIFT_cuda_Tools.cu,
compiled unmodified for sm_100a by oracle/Makefile.  It is a baseline, not an oracle: SURVEY.md section 0 lists
why its output is not trusted (it mirrors every volume over PCIe and clobbers blur inputs).  This script writes
the MNI phantom as NIfTI, runs the CLI with and without -d0 and reports wall-clock seconds, the sum of the
reference's own '#<microseconds>' stage lines and the number of feature rows each run wrote.

    python oracle/ref_cuda_baseline.py [--reps 3] [--shape 182,218,182]
"""
import argparse
import importlib.util
import json
import os
import re
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def run(binary, args, cwd, timeout):
    t = time.perf_counter()
    try:
        p = subprocess.run([binary] + args, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"status": "timeout after %d s" % timeout}
    dt = time.perf_counter() - t
    stage_us = [int(m) for m in re.findall(r"^#(\d+)\s*$", p.stdout, flags=re.M)]
    rows = None
    out = os.path.join(cwd, args[-1])
    if os.path.exists(out):
        for line in open(out, errors="ignore"):
            if line.startswith("Features:"):
                rows = int(line.split()[1])
                break
    return {"status": "exit %d" % p.returncode, "wall_s": dt, "stage_sum_s": sum(stage_us) * 1e-6, "stages": len(stage_us),
            "rows": rows, "stderr_tail": p.stderr[-300:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--shape", default="182,218,182")
    ap.add_argument("--timeout", type=int, default=120)
    a = ap.parse_args()
    spec = importlib.util.spec_from_file_location("s3d_phantom", os.path.join(ROOT, "3d_sift_cuda_b200", "phantom.py"))
    ph = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ph)
    shape = tuple(int(v) for v in a.shape.split(","))
    vol = ph.brain_phantom(shape, 1, 400)
    cuda_bin = os.path.join(HERE, "_ref", "featExtract_ref_cuda")
    cpu_bin = os.path.join(HERE, "_ref", "featExtract_ref")
    res = {"shape_xyz": list(shape), "binary": os.path.relpath(cuda_bin, ROOT)}
    if not os.path.exists(cuda_bin):
        res["unavailable"] = "oracle/_ref/featExtract_ref_cuda not built (needs /root/reference at build time)"
        print(json.dumps(res))
        return 0
    with tempfile.TemporaryDirectory() as d:
        ph.write_nifti(os.path.join(d, "in.nii"), vol)
        res["cpu"] = run(cpu_bin, ["in.nii", "cpu.key"], d, a.timeout * 3) if os.path.exists(cpu_bin) else None
        runs = [run(cuda_bin, ["-d0", "in.nii", "d0.key"], d, a.timeout) for _ in range(a.reps)]
        res["d0_runs"] = runs
        ok = [r for r in runs if r.get("status") == "exit 0"]
        if ok:
            res["d0_best_wall_s"] = min(r["wall_s"] for r in ok)
            res["d0_best_stage_sum_s"] = min(r["stage_sum_s"] for r in ok)
            res["d0_volumes_per_s_stage_sum"] = 1.0 / res["d0_best_stage_sum_s"] if res["d0_best_stage_sum_s"] > 0 else None
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
