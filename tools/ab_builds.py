"""Build A/B variants of libbdl.so (different -D knobs) in seconds and time them on the GPU box.

    python tools/ab_builds.py build  NAME="-DFLAG=1 -DOTHER=2" NAME2="..."     # here (nvcc cross-compiles)
    python tools/ab_builds.py run [--only REGEX] [--steps N] [--rounds R]      # on the GPU box (gpurun)

Builds use -DBDL_AB_SLIM (bdl_step.cu: only the SGHMC / Adam-cSGHMC Philox + reciprocal instantiations) and land in
tools/_ab/NAME/libbdl.so (git-ignored, travels with the gpurun snapshot); `run` launches tools/ab_block.py once per
build and round with BDL_LIB_PATH pointing at it, so clock drift under the power cap hits every build alike.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "tools", "_ab")
sys.path.insert(0, ROOT)


def build(specs):
    from bayesdll_b200 import build as b
    nvcc = b._nvcc()
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(AB, name)
        os.makedirs(out, exist_ok=True)
        objs, procs = [], []
        for src in b.SOURCES:
            obj = os.path.join(out, src.replace(".cu", ".o"))
            objs.append(obj)
            cmd = [nvcc, *b.NVCC_FLAGS, "-DBDL_AB_SLIM", *flags.split(), "-c", os.path.join(b.CSRC, src), "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for src, p in procs:
            o, _ = p.communicate()
            if p.returncode:
                raise SystemExit(f"{name}: nvcc failed on {src}\n{o}")
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o",
                               os.path.join(out, "libbdl.so"), *objs, "-lcudart"])
        with open(os.path.join(out, "flags.txt"), "w") as f:
            f.write(flags + "\n")
        for o in objs:
            os.remove(o)
        print(f"built {name}: {flags}")


def run(argv):
    rounds = 2
    script = os.path.join(ROOT, "tools", "ab_block.py")
    if "--script" in argv:                                 # e.g. --script tools/ab_draw.py (takes no further arguments)
        i = argv.index("--script")
        script = os.path.join(ROOT, argv[i + 1])
        del argv[i:i + 2]
    if "--rounds" in argv:
        i = argv.index("--rounds")
        rounds = int(argv[i + 1])
        del argv[i:i + 2]
    names = sorted(d for d in os.listdir(AB) if os.path.exists(os.path.join(AB, d, "libbdl.so")))
    for r in range(rounds):
        for name in names:
            flags = open(os.path.join(AB, name, "flags.txt")).read().strip()
            print(f"=== round {r} build {name} [{flags}]", flush=True)
            env = dict(os.environ, BDL_LIB_PATH=os.path.join(AB, name, "libbdl.so"))
            extra = ["--rounds", "1", *argv] if script.endswith("ab_block.py") else argv
            subprocess.call([sys.executable, script, *extra], env=env)


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] not in ("build", "run"):
        raise SystemExit(__doc__)
    build(sys.argv[2:]) if sys.argv[1] == "build" else run(sys.argv[2:])
