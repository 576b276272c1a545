"""ctypes binding of include/mclip.h (the C ABI).  No torch types cross this boundary: tensors are passed as
`data_ptr()` integers and the stream as `torch.cuda.current_stream().cuda_stream`.

There is NO CPU fallback: if the library is missing or the device is not a B200, calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCLIP_LIB") or os.path.join(_HERE, "lib", "libmclip_b200.so")     # MCLIP_LIB: A/B builds of the same ABI

MAX_TENSORS, MAX_PAIRS = 4, 8


class MclipError(RuntimeError):
    pass


class LossArgs(C.Structure):
    _fields_ = [
        ("world", C.c_int), ("rank", C.c_int), ("batch", C.c_int), ("dim", C.c_int),
        ("n_tensors", C.c_int), ("n_pairs", C.c_int),
        ("local", C.c_void_p * MAX_TENSORS),
        ("grad", C.c_void_p * MAX_TENSORS),
        ("pair_a", C.c_int * MAX_PAIRS), ("pair_b", C.c_int * MAX_PAIRS),
        ("w_row", C.c_float * MAX_PAIRS), ("w_col", C.c_float * MAX_PAIRS), ("label_smoothing", C.c_float * MAX_PAIRS),
        ("logit_scale", C.c_float),
        ("logit_scale_dev", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong),
        ("out", C.c_void_p),
        ("gathered", C.c_void_p * MAX_TENSORS),
        ("peer_gathered", C.c_void_p),
        ("peer_flags", C.c_void_p),
        ("my_flags", C.c_void_p),
        ("epoch", C.c_longlong),
        ("status", C.c_void_p), ("peer_timeout_s", C.c_double),
        ("lse_all", C.c_void_p), ("peer_lse", C.c_void_p),
    ]


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_longlong), ("a_batch_stride", C.c_longlong),
        ("b", C.c_void_p), ("ldb", C.c_longlong), ("b_batch_stride", C.c_longlong),
        ("d", C.c_void_p), ("ldd", C.c_longlong), ("d_batch_stride", C.c_longlong),
        ("m", C.c_int), ("n", C.c_int), ("k", C.c_int), ("batches", C.c_int),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_longlong), ("r_batch_stride", C.c_longlong),
        ("act", C.c_int),
        ("stats", C.c_void_p), ("stat_slots", C.c_int),
        ("dropmask", C.c_void_p), ("drop_scale", C.c_float),
        ("aux_pre", C.c_void_p), ("ld_aux", C.c_longlong),
        ("a2", C.c_void_p), ("lda2", C.c_longlong), ("a2_batch_stride", C.c_longlong), ("k2", C.c_int),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_longlong),
        ("b", C.c_void_p), ("ldb", C.c_longlong),
        ("out", C.c_void_p), ("ldo", C.c_longlong), ("accumulate", C.c_int),
        ("r", C.c_int), ("i", C.c_int), ("j", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong),
    ]


_lib = None


def lib():
    """Loads the shared library (building it first if nvcc is present and the sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.environ.get("MCLIP_REBUILD") == "1":
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise MclipError(f"{LIB_PATH} is missing: run `python mammo-clip_b200/build.py` (there is no fallback path)")
    L = C.CDLL(LIB_PATH)
    L.mclip_last_error.restype = C.c_char_p
    L.mclip_loss_workspace_bytes.restype = C.c_longlong
    L.mclip_gemm_wgrad_workspace_bytes.restype = C.c_longlong
    L.mclip_colsum_workspace_bytes.restype = C.c_longlong
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mclip_last_error().decode(errors="replace")
        raise MclipError(f"{what} failed ({rc}): {msg}")


# kernels launched per ABI call (for bench.py's `gpu_launches`); default 1
LAUNCHES = {"mclip_gemm_wgrad": 2, "mclip_dwconv_backward": 2, "mclip_se_fc_backward": 2, "mclip_layernorm_backward": 2,
            "mclip_bert_embed_backward": 4, "mclip_colsum": 2, "mclip_bert_attention_backward": 2}


class Profiler:
    """Launch counter + optional CUDA-event timing per ABI entry point (events on torch's current stream, which is the
    stream every kernel is launched on).  Timing is off unless `enable(names)` was called (bench.py roofline leg)."""

    def __init__(self):
        self.counts, self.timed, self.records = {}, None, {}

    def reset(self):
        self.counts, self.records = {}, {}

    def enable(self, names):
        self.timed = set(names) if names is not None else None
        self.records = {}

    def launches(self):
        return sum(LAUNCHES.get(k, 1) * v for k, v in self.counts.items())

    def summary(self):
        """-> {name: (calls, total_ms, total_bytes)} for timed entry points (synchronises)."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = sum(a.elapsed_time(b) for a, b, _ in recs)
            out[name] = (len(recs), ms, sum(n for _, _, n in recs))
        return out


PROF = Profiler()


def call(name, *args, nbytes=0):
    """Invoke C-ABI entry `name(*args, stream)`; raises MclipError on a non-zero return code."""
    import torch
    fn = getattr(lib(), name)
    PROF.counts[name] = PROF.counts.get(name, 0) + 1
    if PROF.timed is not None and name in PROF.timed:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args, stream_ptr())
        b.record()
        PROF.records.setdefault(name, []).append((a, b, nbytes))
    else:
        rc = fn(*args, stream_ptr())
    check(rc, name)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


class DwconvArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int), ("ho", C.c_int), ("wo", C.c_int), ("k", C.c_int), ("stride", C.c_int),
        ("pad_left", C.c_int), ("pad_right", C.c_int), ("pad_top", C.c_int), ("pad_bottom", C.c_int),
        ("in_", C.c_void_p), ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("in_act", C.c_int),
        ("weight", C.c_void_p), ("out", C.c_void_p),
        ("stats", C.c_void_p), ("stat_slots", C.c_int),
        ("dy", C.c_void_p), ("dx", C.c_void_p), ("dweight", C.c_void_p), ("accumulate", C.c_int),
        ("dw_partials", C.c_void_p), ("bn_partials", C.c_void_p), ("in_mean", C.c_void_p), ("in_invstd", C.c_void_p),
    ]


class StemArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("ho", C.c_int), ("wo", C.c_int),
        ("pad_left", C.c_int), ("pad_right", C.c_int), ("pad_top", C.c_int), ("pad_bottom", C.c_int),
        ("in_", C.c_void_p), ("stride_n", C.c_longlong), ("stride_c", C.c_longlong), ("stride_h", C.c_longlong), ("stride_w", C.c_longlong),
        ("out", C.c_void_p),
        ("in_channels", C.c_int), ("in_dtype", C.c_int), ("norm_lut", C.c_void_p),
    ]


class Bn0FoldArgs(C.Structure):
    _fields_ = [
        ("cexp", C.c_int), ("cin", C.c_int), ("k1pad", C.c_int), ("ldw", C.c_longlong), ("count", C.c_double),
        ("we", C.c_void_p), ("scale", C.c_void_p), ("invstd", C.c_void_p), ("c1", C.c_void_p), ("c2", C.c_void_p),
        ("wcat", C.c_void_p), ("twe", C.c_void_p), ("bias", C.c_void_p),
        ("g", C.c_void_p), ("sumx", C.c_void_p), ("gc", C.c_void_p), ("dwe", C.c_void_p), ("q", C.c_void_p),
    ]


class BnArgs(C.Structure):
    _fields_ = [
        ("c", C.c_int), ("slots", C.c_int), ("training", C.c_int), ("count", C.c_longlong),
        ("partials", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
        ("momentum", C.c_float), ("eps", C.c_float),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
    ]


class EwArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("hw", C.c_int), ("c", C.c_int), ("act", C.c_int), ("chunks", C.c_int),
        ("y", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("rowscale", C.c_void_p), ("residual", C.c_void_p),
        ("out", C.c_void_p), ("pool_partials", C.c_void_p),
    ]


class SeArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("hw", C.c_int), ("c", C.c_int), ("cse", C.c_int), ("chunks", C.c_int), ("accumulate", C.c_int),
        ("pool_partials", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("pooled", C.c_void_p), ("z1", C.c_void_p), ("gate", C.c_void_p),
        ("dgate_partials", C.c_void_p), ("dgate_chunk_stride", C.c_int), ("dz2", C.c_void_p), ("dz1", C.c_void_p), ("dpool", C.c_void_p),
        ("dw1", C.c_void_p), ("db1", C.c_void_p), ("dw2", C.c_void_p), ("db2", C.c_void_p),
    ]


class EwBwdArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("hw", C.c_int), ("c", C.c_int), ("act", C.c_int), ("mode", C.c_int), ("dv_given", C.c_int), ("chunks", C.c_int),
        ("y", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("du", C.c_void_p), ("dvec", C.c_void_p), ("gate", C.c_void_p), ("dpool", C.c_void_p), ("rowscale", C.c_void_p),
        ("mean", C.c_void_p), ("invstd", C.c_void_p), ("c1", C.c_void_p), ("c2", C.c_void_p),
        ("partials", C.c_void_p), ("out", C.c_void_p),
    ]


class PrepEntry(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dst_t", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("dst_ld", C.c_int), ("dst_t_ld", C.c_int)]


class BertEmbedArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("seq_len", C.c_int), ("hidden", C.c_int), ("vocab", C.c_int), ("max_positions", C.c_int),
        ("input_ids", C.c_void_p), ("token_type_ids", C.c_void_p),
        ("word", C.c_void_p), ("pos", C.c_void_p), ("type", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("dropmask", C.c_void_p), ("drop_scale", C.c_float),
        ("out", C.c_void_p),
    ]


class BertEmbedBwdArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("seq_len", C.c_int), ("hidden", C.c_int), ("vocab", C.c_int), ("max_positions", C.c_int),
        ("n_types", C.c_int), ("slots", C.c_int), ("accumulate", C.c_int),
        ("input_ids", C.c_void_p), ("token_type_ids", C.c_void_p),
        ("word", C.c_void_p), ("pos", C.c_void_p), ("type", C.c_void_p), ("gamma", C.c_void_p), ("eps", C.c_float),
        ("dropmask", C.c_void_p), ("drop_scale", C.c_float),
        ("dout", C.c_void_p),
        ("dv", C.c_void_p), ("partials", C.c_void_p),
        ("dword", C.c_void_p), ("dpos", C.c_void_p), ("dtype", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
    ]
