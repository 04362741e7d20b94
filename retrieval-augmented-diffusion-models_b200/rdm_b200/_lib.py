"""ctypes binding of librdm_b200.so (the C ABI declared in include/rdm_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
PyTorch is used only for device memory, streams and torch.distributed plumbing.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDM_B200_LIB") or os.path.join(_HERE, "librdm_b200.so")     # RDM_B200_LIB: developer A/B builds (tools/ab_build.sh)
CSRC_DIR = os.path.join(os.path.dirname(_HERE), "csrc")

_lib = None

c_void_p, c_int, c_int32, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64
c_float, c_double, c_char_p, c_ulonglong = ctypes.c_float, ctypes.c_double, ctypes.c_char_p, ctypes.c_ulonglong

# name -> (restype, argtypes); kept in sync with include/rdm_b200.h (tests/test_abi.py checks both ways)
SIGNATURES = {
    "rdm_last_error": (c_char_p, []),
    "rdm_abi_version": (c_int, []),
    "rdm_launch_count": (c_ulonglong, []),
    "rdm_knn_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int64, c_int32, c_int32, c_int32, c_int64, c_int32]),
    "rdm_knn_destroy": (None, [c_void_p]),
    "rdm_knn_size": (c_int64, [c_void_p]),
    "rdm_knn_get_inv_norms": (c_int, [c_void_p, c_void_p, c_void_p]),
    "rdm_knn_search": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_knn_normalize": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "rdm_knn_search_raw": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_knn_merge": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "rdm_knn_gather": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "rdm_unet_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int32]),
    "rdm_unet_destroy": (None, [c_void_p]),
    "rdm_unet_num_params": (c_int64, [c_void_p]),
    "rdm_unet_param_name": (c_char_p, [c_void_p, c_int64]),
    "rdm_unet_param_numel": (c_int64, [c_void_p, c_char_p]),
    "rdm_unet_load": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "rdm_unet_missing": (c_int64, [c_void_p]),
    "rdm_unet_set_mode": (c_int, [c_void_p, c_int32]),
    "rdm_unet_set_debug": (c_int, [c_void_p, c_int32]),
    "rdm_unet_debug_log": (c_char_p, [c_void_p]),
    "rdm_unet_set_context": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "rdm_unet_forward": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "rdm_unet_set_graph": (c_int, [c_void_p, c_int32]),
    "rdm_unet_set_chains": (c_int, [c_void_p, c_int32]),
    "rdm_unet_set_ablation": (c_int, [c_void_p, c_int32]),
    "rdm_unet_profile_forward": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "rdm_unet_profile_text": (c_char_p, [c_void_p]),
    "rdm_ddim_sample": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    "rdm_vqdec_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int32]),
    "rdm_vqdec_decode": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "rdm_clip_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int32]),
    "rdm_clip_destroy": (None, [c_void_p]),
    "rdm_clip_num_params": (c_int64, [c_void_p]),
    "rdm_clip_param_name": (c_char_p, [c_void_p, c_int64]),
    "rdm_clip_param_numel": (c_int64, [c_void_p, c_char_p]),
    "rdm_clip_load": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "rdm_clip_missing": (c_int64, [c_void_p]),
    "rdm_clip_set_mode": (c_int, [c_void_p, c_int32]),
    "rdm_clip_encode_text": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "rdm_clip_encode_image": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "rdm_clip_preprocess": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "rdm_rarm_create": (c_int, [ctypes.POINTER(c_void_p), c_void_p, c_int32]),
    "rdm_rarm_destroy": (None, [c_void_p]),
    "rdm_rarm_num_params": (c_int64, [c_void_p]),
    "rdm_rarm_param_name": (c_char_p, [c_void_p, c_int64]),
    "rdm_rarm_param_numel": (c_int64, [c_void_p, c_char_p]),
    "rdm_rarm_load": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "rdm_rarm_missing": (c_int64, [c_void_p]),
    "rdm_rarm_set_mode": (c_int, [c_void_p, c_int32]),
    "rdm_rarm_set_graph": (c_int, [c_void_p, c_int32]),
    "rdm_rarm_set_context": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "rdm_rarm_forward_token": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "rdm_rarm_sample_step": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rdm_rarm_sample": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_int32, c_float, c_void_p, c_void_p]),
    "rdm_ddim_step": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
}


def build(verbose=False):
    """Compile the CUDA sources in csrc/ for sm_100a into librdm_b200.so (in-tree)."""
    import subprocess
    out = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building librdm_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the product has no CPU or PyTorch fallback)")
        _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


def bind(L, names=None):
    """Applies the SIGNATURES table to a loaded library (all entry points, or the given subset)."""
    for name in (SIGNATURES if names is None else names):
        fn = getattr(L, name)
        fn.restype, fn.argtypes = SIGNATURES[name]
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().rdm_last_error()
        raise RuntimeError(f"librdm_b200 {what} failed ({rc}): {msg.decode() if msg else ''}")


def stream_ptr(device=None):
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def resolve_device(device):
    """torch.device with an explicit CUDA index; anything else is an error (the library has no CPU path)."""
    import torch
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"librdm_b200 needs a CUDA device, got '{device}' (there is no CPU path)")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def device_ctx(device):
    """Context manager that makes `device` current for the duration of a library call."""
    import torch
    return torch.cuda.device(device)


def ptr(t):
    """Device/host pointer of a contiguous tensor (or None)."""
    if t is None:
        return c_void_p(0)
    assert t.is_contiguous(), "librdm_b200 needs contiguous tensors"
    return c_void_p(t.data_ptr())


def launch_count():
    return int(lib().rdm_launch_count())
