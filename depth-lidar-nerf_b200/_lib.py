"""Loader for the C-ABI library (include/dlnerf_b200.h) and ctypes mirrors of its structs.

There is NO CPU fallback: if ``libdlnerf_b200.so`` has not been built
(``python -c "import __graft_entry__ as g; g.build()"``) every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("DLN_SO_PATH") or os.path.join(_HERE, "libdlnerf_b200.so")   # override: A/B builds of the kernels
SOURCES = ["render_kernels.cu", "mlp_kernels.cu", "mlp_chain2.cu", "optim_kernels.cu", "semantic_kernels.cu", "raygen_kernels.cu"]

MAX_STEPS = 12
MAX_KSLABS = 6
SLAB_BYTES = 16384
TILE_ROWS = 128
SEM_MAX_CLASSES = 32
WGRAD_PARTIAL_FLOATS = 256 * 256 + 256

EPI_RELU, EPI_RELU_SIGMA, EPI_LINEAR, EPI_RELU_RGB, EPI_RELU_OUT = 0, 1, 2, 3, 4
EPI_BWD_COPY, EPI_BWD_MASK, EPI_BWD_MASK_SIGMA = 8, 9, 10


class ChainStep(C.Structure):
    _fields_ = [("w_off", C.c_uint32), ("bias_off", C.c_uint32), ("head_off", C.c_uint32),
                ("head_bias_off", C.c_uint32), ("n_out", C.c_uint16), ("nk", C.c_uint8), ("epi", C.c_uint8),
                ("kslab", C.c_uint8 * MAX_KSLABS), ("kcnt", C.c_uint8 * MAX_KSLABS),
                ("stash_slot", C.c_int16), ("mask_slot", C.c_int16), ("n_heads", C.c_uint8),
                ("n_valid32", C.c_uint8), ("pad_", C.c_uint8 * 2)]


class ChainProgram(C.Structure):
    _fields_ = [("n_steps", C.c_int32), ("backward", C.c_int32), ("use_viewdirs", C.c_int32),
                ("out_ch", C.c_int32), ("L_pts", C.c_int32), ("L_dir", C.c_int32), ("stash_slots", C.c_int32),
                ("mask_slots", C.c_int32), ("pro_head_off", C.c_int32), ("pro_mask_slot", C.c_int32),
                ("pro_slot", C.c_int32), ("reload_step", C.c_int32), ("pro_valid", C.c_int32),
                ("steps", ChainStep * MAX_STEPS)]


class ChainArgs(C.Structure):
    _fields_ = [("P", C.c_longlong), ("rays", C.c_void_p), ("ray_stride", C.c_int32), ("vd_col", C.c_int32),
                ("z", C.c_void_p), ("S", C.c_int32), ("x", C.c_void_p), ("x_ld", C.c_int32),
                ("wblob", C.c_void_p), ("fblob", C.c_void_p), ("out", C.c_void_p), ("d_out", C.c_void_p),
                ("stash", C.c_void_p), ("masks", C.c_void_p), ("trace", C.c_void_p),
                ("sem_g", C.c_void_p), ("sem_g_div", C.c_int32), ("z_lindisp", C.c_int32),
                ("z_gen", C.c_void_p), ("z_rng_state", C.c_void_p), ("z_rng_offset", C.c_ulonglong)]


class WgradItem(C.Structure):
    _fields_ = [("a_bwd_stash", C.c_int32), ("a_slot", C.c_int32), ("a_nslab", C.c_int32),
                ("b_from_bwd", C.c_int32), ("b_slot", C.c_int32), ("b_nslab", C.c_int32),
                ("dw_off", C.c_int64), ("ld", C.c_int32), ("col_off", C.c_int32), ("n_cols", C.c_int32),
                ("row_off", C.c_int32), ("n_rows", C.c_int32), ("db_off", C.c_int64),
                ("db_col_off", C.c_int32), ("db_n", C.c_int32)]


class PackJob(C.Structure):
    _fields_ = [("src_off", C.c_int64), ("ld", C.c_int32), ("row0", C.c_int32), ("col0", C.c_int32),
                ("n_valid", C.c_int32), ("k_valid", C.c_int32), ("transposed", C.c_int32),
                ("n_rows", C.c_int32), ("dst_off", C.c_uint32)]


class SemOffsets(C.Structure):
    _fields_ = [("w_f", C.c_int64), ("b_f", C.c_int64), ("w_s1", C.c_int64), ("b_s1", C.c_int64),
                ("w_s2", C.c_int64), ("b_s2", C.c_int64), ("A", C.c_int64), ("a", C.c_int64),
                ("Sw", C.c_int64), ("sc", C.c_int64), ("K", C.c_int32), ("pad_", C.c_int32)]


_P, _I, _F, _LL, _D, _U64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_double, C.c_ulonglong

# name -> argtypes; every function returns int.  Must list EVERY symbol include/dlnerf_b200.h declares
# (tests/test_abi.py parses the header and compares).
SIGNATURES = {
    "dln_pack_rays": [_P, _P, _I, _I, _I, _I, _D, _F, _F, _F, _I, _P, _P],
    "dln_stratified_z": [_P, _I, _P, _P, _I, _I, _I, _P],
    "dln_rng_advance": [_P, _U64, _P],
    "dln_rng_fill": [_P, _U64, _I, _P, _LL, _I, _P],
    "dln_stratified_z_rng": [_P, _I, _P, _U64, _P, _I, _I, _I, _P],
    "dln_composite_fwd_rng": [_P, _I, _P, _P, _P, _U64, _F, _I, _P, _P, _P, _P, _P, _I, _I, _P],
    "dln_composite_bwd_rng": [_P, _I, _P, _P, _P, _U64, _F, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "dln_composite_bwd_fused_loss_dev": [_P, _I, _P, _P, _P, _P, _U64, _F, _I, _P, _P, _P, _I, _P, _I, _P, _P, _I, _I, _P],
    "dln_sample_pdf_rng": [_P, _I, _I, _P, _I, _I, _P, _U64, _I, _P, _P, _I, _P, _P, _P, _I, _P],
    "dln_composite_resample_fwd": [_P, _I, _P, _P, _P, _P, _U64, _F, _I, _P, _P, _P, _P, _P, _P, _P, _U64, _I, _P, _P, _I,
                                   _I, _P],
    "dln_posenc": [_P, _P, _LL, _I, _P],
    "dln_composite_fwd": [_P, _I, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P, _I, _I, _P],
    "dln_composite_bwd": [_P, _I, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "dln_composite_bwd_fused_loss": [_P, _I, _P, _P, _P, _F, _I, _P, _P, _P, _I, _F, _F, _I, _F, _P, _P, _I, _I, _P],
    "dln_sample_pdf": [_P, _I, _I, _P, _I, _I, _P, _I, _P, _P, _I, _P, _P, _P, _I, _P],
    "dln_searchsorted": [_P, _I, _I, _P, _I, _I, _P, _I, _P],
    "dln_mlp_chain": [C.POINTER(ChainProgram), C.POINTER(ChainArgs), _I, _P],
    "dln_mlp_wgrad": [_P, _I, _I, _P, _I, _P, _I, _LL, _P, _P, _P],
    "dln_mlp_pack_weights": [_P, _P, _I, _P, _P],
    "dln_inv_depth_smooth_fwd": [_P, _P, _I, _I, _I, _P, _P],
    "dln_inv_depth_smooth_bwd": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "dln_mlp_fold": [_P, _LL, _I, _LL, _LL, _LL, _LL, _LL, _P],
    "dln_mlp_unfold_grads": [_P, _P, _LL, _I, _LL, _LL, _LL, _LL, _LL, _P],
    "dln_sem_fold": [_P, C.POINTER(SemOffsets), _P],
    "dln_sem_unfold_grads": [_P, _P, C.POINTER(SemOffsets), _P],
    "dln_sem_head_fwd": [_P, _I, _I, _LL, _I, _P, C.POINTER(SemOffsets), _P, _P, _I, _P],
    "dln_sem_head_bwd": [_P, _I, _P, _LL, _I, _P, _P, C.POINTER(SemOffsets), _P, _P],
    "dln_sample_sum": [_P, _I, _I, _I, _I, _P, _P],
    "dln_sample_sum_bwd": [_P, _I, _I, _I, _I, _P, _P],
    "dln_sem_ce_loss": [_P, _I, _P, _I, _I, _I, _F, _P, _P, _P],
    "dln_adam_step": [_P, _P, _P, _P, _LL, _D, _D, _D, _D, _I, _F, _P],
    "dln_gen_rays": [_P, _I, _I, _I, _D, _P, _P, _LL, _P],
    "dln_gen_rays_by_coord": [_P, _P, _LL, _I, _I, _D, _I, _P, _P, _LL, _P],
    "dln_gen_rays_patch": [_P, _I, _I, _D, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P],
    "dln_abi_sizes": [C.POINTER(C.c_int)],
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(_HERE, "build")


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU).  Every
    source becomes its own object (compiled in parallel, rebuilt only when it or a header changed) and the objects
    are linked into libdlnerf_b200.so."""
    import glob
    from concurrent.futures import ThreadPoolExecutor
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "dlnerf_b200.h")]
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    if not force and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= max([hdr_time] + [os.path.getmtime(x) for x in srcs]):
        return SO_PATH          # up to date (the objects do not travel to the GPU box, the library does)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("DLN_NVCC_EXTRA", "").split()
    objs, todo = [], []
    for src in srcs:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            todo.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    if todo:
        with ThreadPoolExecutor(max_workers=len(todo)) as ex:
            results = list(ex.map(compile_one, todo))
        for src, res in results:
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
            if verbose:
                print(res.stderr)
    if todo or not os.path.exists(SO_PATH) or any(os.path.getmtime(SO_PATH) < os.path.getmtime(o) for o in objs):
        res = subprocess.run([nvcc, "-shared", "-o", SO_PATH] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded library; raises (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "dlnerf_b200: %s is missing; there is no CPU fallback. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'`." % SO_PATH)
        l = C.CDLL(SO_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        l.dln_build_info.restype = C.c_char_p
        l.dln_build_info.argtypes = []
        _lib = l
    return _lib


class DlnError(RuntimeError):
    pass


# Launch accounting (bench.py): LAUNCHES counts kernel-launching ABI calls; when TRACE is a list every call
# is bracketed by CUDA events recorded on the launching (current) stream and appended as (tag, ev0, ev1).
LAUNCHES = 0
TRACE = None


def call(name: str, *args, tag: str = None) -> None:
    """Invoke one ABI entry point, raise on a non-zero return code."""
    global LAUNCHES
    fn = getattr(lib(), name)
    LAUNCHES += 1
    if TRACE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        TRACE.append((tag or name, e0, e1))
    else:
        rc = fn(*args)
    check(rc, tag or name)


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc == -1:
        raise DlnError("%s: invalid argument (shape / alignment / null pointer)" % what)
    raise DlnError("%s: CUDA error %d" % (what, rc))
