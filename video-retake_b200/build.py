"""Build librtk_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build() and `make`.

The build is gated on a HASH of the sources, not on file times: ``source_id()`` is the first 16 hex digits of the SHA-256
over csrc/*.cu, csrc/*.cuh and include/rtk_b200.h (names and contents, sorted).  It is compiled into the library
(``rtk_build_id()``) and written next to it (``lib/BUILD_ID``); ``build()`` recompiles whenever the id of the sources
differs from the id of the library, and the Python binding (``retake/_native.py``) refuses to load a library whose id does
not match the sources lying next to it - so a test or bench record made from this tree always ran these sources.
After linking, ``lib/SASS_SUMMARY.json`` lists per object the opcodes that prove the Blackwell paths (UTCHMMA, UTMALDG,
LDTM, UBLKCP ...), from ``cuobjdump -sass``."""
import hashlib
import json
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(HERE, "..", "include", "rtk_b200.h")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "librtk_b200.so")
ID_FILE = os.path.join(OUT_DIR, "BUILD_ID")
SOURCES = ["dpselect.cu", "mallm.cu", "pivot_score.cu", "pivot_misc.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUOBJDUMP = os.path.join(os.path.dirname(NVCC), "cuobjdump")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-DNDEBUG"]
SASS_OPS = ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "MUFU.EX2", "F2FP", "FHADD", "FHFMA",
            "FFMA2", "FMUL2", "FADD2", "SYNCS", "HMMA")


def source_id() -> str:
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    for name, path in [(f, os.path.join(CSRC, f)) for f in files] + [("rtk_b200.h", HEADER)]:
        h.update(name.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()[:16]


def built_id():
    try:
        return open(ID_FILE).read().strip() if os.path.exists(LIB) else None
    except OSError:
        return None


def sass_summary(objs):
    out = {}
    for obj in objs:
        r = subprocess.run([CUOBJDUMP, "-sass", obj], capture_output=True, text=True)
        if r.returncode != 0:
            continue
        counts = {}
        for op in SASS_OPS:
            n = len(re.findall(r"\b" + re.escape(op), r.stdout))
            if n:
                counts[op] = n
        out[os.path.basename(obj)] = counts
    return out


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    sid = source_id()
    if not force and built_id() == sid:
        return LIB

    def compile_one(src):
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, f'-DRTK_BUILD_ID="{sid}"', "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    if os.path.exists(ID_FILE):
        os.remove(ID_FILE)                       # a failed build must not leave an id that vouches for an old library
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    with open(os.path.join(OUT_DIR, "SASS_SUMMARY.json"), "w") as f:
        json.dump({"build_id": sid, "opcodes": sass_summary(objs)}, f, indent=1, sort_keys=True)
    with open(ID_FILE, "w") as f:
        f.write(sid + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv), source_id())
