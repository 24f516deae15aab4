"""Build libmst_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m music_mixing_style_transfer_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles here without a GPU.
"""
import argparse
import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libmst_b200.so")
STAMP = LIB_PATH + ".stamp"
SOURCES = ["api.cu", "encoder.cu", "enc_umma.cu", "tcn.cu", "tcn_f8.cu", "fx2.cu", "fxnorm.cu", "spectral.cu", "reverb.cu", "pcm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--cudart", "static"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    # names relative to the package, so a shipped .so is still "fresh" when the repo is mounted at another path
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, "mst_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _digest()


def build_variant(tag: str, defines, csrc: str = None) -> str:
    """Development builds with extra -D switches (ablation studies) or from another source directory (A/B against an older
    commit: `git archive <rev> music_mixing_style_transfer_b200/csrc include | tar -x -C /tmp/rev`): build/<tag>/libmst_b200.so,
    never the product path."""
    csrc = CSRC if csrc is None else csrc
    out_dir = os.path.join(PKG_DIR, "build", tag)
    os.makedirs(out_dir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(out_dir, src.replace(".cu", ".o"))
        if not os.path.exists(os.path.join(csrc, src)):
            continue
        procs.append(subprocess.Popen([_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", INCLUDE, "-c",
                                       os.path.join(csrc, src), "-o", obj], stdout=subprocess.DEVNULL,
                                      stderr=subprocess.DEVNULL))
        objs.append(obj)
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("nvcc failed")
    lib = os.path.join(out_dir, "libmst_b200.so")
    subprocess.check_call([_nvcc(), "-shared", *NVCC_FLAGS, "-o", lib, *objs])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed; see output above")
    link = [_nvcc(), "-shared", *NVCC_FLAGS, "-o", LIB_PATH, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", nargs="+", metavar=("TAG", "DEFINE"), help="development build: TAG then -D defines")
    ap.add_argument("--csrc", default=None, help="with --variant: compile the sources of this directory instead")
    a = ap.parse_args()
    print(build_variant(a.variant[0], a.variant[1:], a.csrc) if a.variant else build(a.force, a.verbose))
