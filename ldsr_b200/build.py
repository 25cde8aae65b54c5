"""Builds ldsr_b200/libldsr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per padded input width PQ (kernels_inst.cu -DLDSR_PQ=n) plus the host/ABI
unit; objects are compiled in parallel and linked with a static CUDA runtime so the library has
no dependency beyond the driver.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libldsr_b200.so")
PQ_LIST = (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 24, 32)  # keep in sync with LDSR_PQ_LIST in ldsr_abi.cu
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SPLIT_TU_FROM_PQ = 12  # widths from here on are compiled as 6 translation units each
WIDE_FROM_PQ = 5       # keep in sync with WIDE_MIN_PQ (kernel_table.h): these widths have the wide-input kernel
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(HERE, "..", "include", "ldsr_b200.h"))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force=False, verbose=False, ptxas_info=False):
    """Env (development only): LDSR_PQ_LIST="3,10" builds just those widths; LDSR_SEG=4 overrides
    the checkpoint segment length; LDSR_NVCC_EXTRA adds nvcc flags; LDSR_BUILD_TAG=x writes
    variants/lib_x.so (loaded with LDSR_SO=...) instead of the product library."""
    global PQ_LIST, OBJ, SO
    if os.environ.get("LDSR_BUILD_TAG"):  # side-by-side experimental builds: variants/lib_<tag>.so
        tag = os.environ["LDSR_BUILD_TAG"]
        OBJ = os.path.join(HERE, "build_" + tag)
        os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
        SO = os.path.join(HERE, "variants", "lib_%s.so" % tag)
    if os.environ.get("LDSR_PQ_LIST"):
        PQ_LIST = tuple(int(x) for x in os.environ["LDSR_PQ_LIST"].split(","))
    deps = _sources()
    # per-object dependencies, so that touching the ABI or the scan kernels does not recompile the
    # thirteen per-width EM translation units (minutes)
    c = lambda *names: [os.path.join(CSRC, n) for n in names]
    em_deps = c("kernels_inst.cu", "kernel_table.h", "em_kernel.cuh", "em_split_kernel.cuh", "em_wide_kernel.cuh",
                "em_scan_kernel.cuh", "aux_kernels.cuh", "lds_math.cuh", "common.cuh")
    scan_deps = c("scan_inst.cu", "scan_kernels.cuh", "common.cuh")
    obj_deps = {}
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    if os.environ.get("LDSR_SEG"):
        extra += ["-DLDSR_SEG=" + os.environ["LDSR_SEG"]]
    extra += os.environ.get("LDSR_NVCC_EXTRA", "").split()
    # widest first (longest jobs first); a wide PQ is cut into 5 parts (see kernels_inst.cu) so that no
    # single nvcc process is the critical path of the build
    for pq in sorted(PQ_LIST, reverse=True):
        # None: everything in one unit; "main": everything but the wide-input kernel (its own unit, part 5)
        parts = [None] if pq < WIDE_FROM_PQ else (["main", 5] if pq < SPLIT_TU_FROM_PQ else [1, 2, 3, 4, 5, 0])
        for part in parts:
            o = os.path.join(OBJ, "kernels_pq%d%s.o" % (pq, "" if part is None else "_part%s" % part))
            sel = [] if part is None else (["-DLDSR_NO_PART=5"] if part == "main" else ["-DLDSR_PART=%d" % part])
            cmd = [NVCC] + FLAGS + extra + ["-DLDSR_PQ=%d" % pq] + sel
            jobs.append((o, cmd + ["-c", os.path.join(CSRC, "kernels_inst.cu"), "-o", o]))
            obj_deps[o] = em_deps
    o_scan = os.path.join(OBJ, "scan_inst.o")
    jobs.append((o_scan, [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, "scan_inst.cu"), "-o", o_scan]))
    obj_deps[o_scan] = scan_deps
    o_abi = os.path.join(OBJ, "ldsr_abi.o")
    jobs.append((o_abi, [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, "ldsr_abi.cu"), "-o", o_abi]))
    todo = [(o, cmd) for o, cmd in jobs if force or _stale(o, obj_deps.get(o, deps))]
    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        for out in ex.map(lambda j: _run(j[1]), todo):
            logs.append(out)
    objs = [o for o, _ in jobs]
    # staleness is decided per object (above); the library is relinked when an object is newer than it
    if force or todo or _stale(SO, objs):
        logs.append(_run([NVCC, "-shared", "-o", SO] + objs + ["-cudart", "static", "-lpthread"]))
    if verbose:
        sys.stderr.write("".join(logs))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
