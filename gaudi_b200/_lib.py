"""ctypes binding of ``csrc/libgaudi_b200.so`` (the C ABI declared in ``include/gaudi_b200.h``).

There is no fallback: if the shared library is missing it is built with ``make`` (nvcc, sm_100a) and an
ImportError is raised when that is impossible.  All compute entry points need a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import glob
import hashlib
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libgaudi_b200.so")
# development aid (tools/kernel_lab.py): load an alternative build of the same sources, e.g. compiled with other -D flags
_LIB_OVERRIDE = os.environ.get("GAUDI_B200_LIB")

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_U64 = C.c_ulonglong
_SZ = C.c_size_t

# name -> (restype, argtypes); mirrors include/gaudi_b200.h one to one
SIGNATURES = {
    "gb_abi_version": (_I, []),
    "gb_last_error": (C.c_char_p, []),
    "gb_launch_count": (C.c_longlong, [_I]),
    "gb_denoiser_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _I, _F, _F, _F, C.POINTER(_P), _I, _P]),
    "gb_predictor_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _I, _F, C.POINTER(_P), _I, _P]),
    "gb_net_destroy": (_I, [_P]),
    "gb_net_hidden_padded": (_I, [_P]),
    "gb_tile_pack": (_I, [_P, _I, _P, C.POINTER(_I)]),
    "gb_tile_pack_graphs": (_I, [_P, _I, _I, _P, C.POINTER(_I)]),
    "gb_stage_rows": (_I, []),
    "gb_graph_create": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gb_graph_destroy": (_I, [_P]),
    "gb_denoiser_workspace_bytes": (_SZ, [_P, _P]),
    "gb_predictor_workspace_bytes": (_SZ, [_P, _P, _I]),
    "gb_denoiser_forward": (_I, [_P, _P, _P, _P, _I, _P, _I, _P, _P, _SZ, _P]),
    "gb_predictor_forward": (_I, [_P, _P, _P, _P, _I, _P, _I, _P, _SZ, _P]),
    "gb_predictor_input_grad": (_I, [_P, _P, _P, _I, _P, _P, _SZ, _P]),
    "gb_step_sample": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _U64, _U64, _I, _P, _P]),
    "gb_step_guide": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P]),
    "gb_decode": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _U64, _U64, _F, _F, _F, _P, _P, _P, _P]),
    "gb_cog_fix": (_I, [_P, _P, _P, _F, _I, _I, _P]),
    "gb_noise": (_I, [_P, _P, _I, _I, _I, _F, _U64, _U64, _P]),
    "gb_sample_loop": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _U64, _P, _P, _SZ, _I, _P]),
    "gb_sample_loop_workspace_bytes": (_SZ, [_P, _P, _P]),
    "gb_profile_kernel": (_I, [_P, _P, _I, _I, _P, _SZ, _I, _P]),
    "gb_den_gcl_forward": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "gb_den_equiv_forward": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gb_den_block_forward": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gb_den_egnn_forward": (_I, [_P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gb_pred_layer_forward": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "gb_pred_egnn_forward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    # training step ops
    "gb_linear_scratch_bytes": (_SZ, [_I, _I, _I]),
    "gb_linear": (_I, [_I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "gb_wgrad_scratch_bytes": (_SZ, [_I, _I]),
    "gb_wgrad": (_I, [_I, _I, _I, _P, _I, _P, _I, _P, _I, _I, _P, _P, _SZ, _P]),
    "gb_gemm": (_I, [_I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P]),
    "gb_colsum": (_I, [_P, _I, _I, _I, _P, _P, _I, _P]),
    "gb_rowdot": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "gb_silu_fwd": (_I, [_P, _P, _SZ, _P]),
    "gb_silu_bwd": (_I, [_P, _P, _P, _SZ, _P]),
    "gb_outer_dsilu": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "gb_edge_pre": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "gb_rowcol_reduce": (_I, [_P, _P, _I, _F, _P, _P, _P]),
    "gb_gather_rows": (_I, [_P, _P, _I, _F, _P, _P]),
    "gb_gate_fwd": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "gb_gate_bwd": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "gb_geom_fwd": (_I, [_P, _P, _F, _P, _P, _P]),
    "gb_geom_bwd": (_I, [_P, _P, _F, _P, _P, _P, _P, _P]),
    "gb_coord_fwd": (_I, [_P, _P, _P, _P, _F, _I, _F, _P, _P, _P]),
    "gb_coord_bwd": (_I, [_P, _P, _P, _P, _F, _I, _F, _P, _P, _P, _P]),
    "gb_resmask": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "gb_den_finish_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "gb_den_finish_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "gb_make_zt": (_I, [_P, _P, _P, _P, _P, _P, _F, _F, _F, _I, _I, _I, _P, _P, _P, _P]),
    "gb_check_stability": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _F, _P, _P, _P, _F, _F, _I, _P, _P]),
    "gb_positions2adj": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gb_vlb_loss": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _I, _I, _I, _P, _P, _P]),
    "gb_pool_mean": (_I, [_P, _I, _I, _I, _P, _P]),
    "gb_pool_mean_bwd": (_I, [_P, _I, _I, _I, _P, _P]),
    "gb_train_loss": (_I, [_P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _I, _I, _I, _P, _P, _P]),
    "gb_adamw_state_doubles": (_SZ, [_I]),
    "gb_adamw_scratch_doubles": (_SZ, []),
    "gb_adamw_amsgrad_clip": (_I, [_P, _P, _P, _P, _P, _SZ, _F, _F, _F, _F, _F, _P, _I, _I, _P, _P]),
}


HASH_PATH = LIB_PATH + ".srchash"


def source_hash() -> str:
    """sha256 over the CUDA sources, headers and Makefile the library is built from (content, not mtime: the tree is
    copied between machines)."""
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                   glob.glob(os.path.join(CSRC, "*.h")) + [os.path.join(CSRC, "Makefile"),
                                                            os.path.join(_HERE, "..", "include", "gaudi_b200.h")])
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_stale() -> bool:
    """True when libgaudi_b200.so is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB_PATH):
        return True
    try:
        with open(HASH_PATH) as fh:
            return fh.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``csrc/libgaudi_b200.so`` (nvcc cross-compiles without a GPU)."""
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        raise ImportError("gaudi_b200: libgaudi_b200.so is missing and nvcc is not available to build it")
    env = dict(os.environ)
    if shutil.which("nvcc") is None:
        env["PATH"] = "/usr/local/cuda/bin:" + env.get("PATH", "")
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, env=env, capture_output=not verbose)
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise ImportError("gaudi_b200: building libgaudi_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout)
    with open(HASH_PATH, "w") as fh:
        fh.write(source_hash() + "\n")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if _LIB_OVERRIDE:
            path = _LIB_OVERRIDE
        else:
            if is_stale():                      # missing, or built from older sources (ABI / behaviour mismatch after a pull)
                build()
            path = LIB_PATH
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class GaudiB200Error(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise GaudiB200Error(lib().gb_last_error().decode("utf-8", "replace"))
