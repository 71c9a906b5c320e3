"""development aid: build build/prof/libmjb200_prof.so, the library with the clock64 probes compiled in
(-DMJB_STAGE_PROF: sub-stage cycles of the two halves, read by tools/prof_stages.py;
 -DMJB_PGS4_PROF: per-warp cycles of k_pgs4, read by tools/prof_pgs4.py)"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

defs = [a for a in sys.argv[1:] if a.startswith("-D")] or ["-DMJB_STAGE_PROF"]
only = [a for a in sys.argv[1:] if not a.startswith("-D")]   # rebuild just these units; the others are reused
PER_FILE = {"mjb_kpart1_lean16.cu": ["-DMJB_STAGE_PROF_FN=mjb_debug_stage_prof_p1"],
            "mjb_kpart2_lean16.cu": ["-DMJB_STAGE_PROF_FN=mjb_debug_stage_prof_p2"],
            "mjb_krollout.cu": ["-DMJB_ROLLOUT_PROF"]}
objdir = os.path.join(g.ROOT, "build", "obj", "prof")
os.makedirs(objdir, exist_ok=True)
os.makedirs(os.path.join(g.ROOT, "prof_build"), exist_ok=True)
names = sorted(f for f in os.listdir(g.CSRC) if f.endswith((".cu", ".cc")))
inc = os.path.join(g.REF, "include")


def one(f):
    obj = os.path.join(objdir, os.path.splitext(f)[0] + ".o")
    if only and f not in only and os.path.exists(obj):
        return obj
    subprocess.check_call([g.NVCC] + g.NVCC_FLAGS + defs + PER_FILE.get(f, []) + ["-I", inc, "-I", os.path.join(g.ROOT, "include"), "-c", "-x", "cu",
                           os.path.join(g.CSRC, f), "-o", obj])
    return obj


with ThreadPoolExecutor(max_workers=8) as ex:
    objs = list(ex.map(one, names))
subprocess.check_call([g.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o",
                       os.path.join(g.ROOT, "prof_build", "libmjb200_prof.so")] + objs)
print("built prof_build/libmjb200_prof.so with", defs)
