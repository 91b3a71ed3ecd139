"""A/B harness for kernel experiments (box-to-box variance is ~1 %, softplus up to 5 %: only same-box comparisons count).

    # here (no GPU): build one library per variant next to the repo root; *.so is git-ignored but ships with gpurun
    python tools/ab_bench.py build base: exp1:-DPNDF_EXP_FOO=1 exp2:-DPNDF_EXP_FOO=2
    # on the GPU box: every variant, interleaved twice, same process conditions
    gpurun -- 'python tools/ab_bench.py run lrelu 65536'

A variant is `name:flags` (flags may be empty = the tree as it is).  `run` uses PNDF_LIBRARY (posendf_b200/_lib.py) to point
tools/quick_bench.py at each library in turn."""
import glob, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "posendf_b200", "csrc", "pndf_capi.cu")
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC"]


def build(variants):
    procs = []
    for v in variants:
        name, _, flags = v.partition(":")
        out = os.path.join(ROOT, f"gpurun_ab_{name}.so")
        procs.append((name, subprocess.Popen(NVCC + flags.split() + ["-o", out, SRC])))
    bad = [n for n, p in procs if p.wait() != 0]
    if bad:
        sys.exit(f"build failed: {bad}")
    print("built", [n for n, _ in procs])


def run(args):
    libs = sorted(glob.glob(os.path.join(ROOT, "gpurun_ab_*.so")))
    if not libs:
        sys.exit("no gpurun_ab_*.so: run `python tools/ab_bench.py build ...` first")
    for rep in range(2):
        for lib in libs:
            env = dict(os.environ, PNDF_LIBRARY=lib)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py")] + args, env=env, capture_output=True, text=True)
            last = (r.stdout.strip().splitlines() or [r.stderr.strip()[-200:]])[-1]
            print(f"{os.path.basename(lib):32s} {last}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] not in ("build", "run"):
        sys.exit(__doc__)
    build(sys.argv[2:]) if sys.argv[1] == "build" else run(sys.argv[2:])
