"""ctypes binding of include/b200glm.h.  Fails loudly when the CUDA extension is missing."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb200glm.so")

OK, DOMAIN, INVALID, CUDA = 0, 1, 2, 3
ABI_VERSION = 5          # B200GLM_ABI_VERSION of include/b200glm.h this binding (Desc layout, signatures) was written for
FAMILY = {"bernoulli_logit": 0, "poisson_log": 1, "normal_id": 2, "binomial_logit": 3, "neg_binomial_2_log": 4,
          "ordered_logistic": 5, "categorical_logit": 6}

# every symbol include/b200glm.h declares
SYMBOLS = [
    "b200glm_create", "b200glm_destroy", "b200glm_append_rows", "b200glm_finalize", "b200glm_num_params", "b200glm_log_prob_grad", "b200glm_log_prob",
    "b200glm_set_state", "b200glm_leapfrog", "b200glm_leapfrog_async", "b200glm_grad_async", "b200glm_sync",
    "b200glm_stream", "b200glm_result_device", "b200glm_comm_unique_id", "b200glm_comm_init",
    "b200glm_batch_reserve", "b200glm_log_prob_grad_batched", "b200glm_set_state_batched",
    "b200glm_leapfrog_batched", "b200glm_leapfrog_batched_async", "b200glm_batch_sync", "b200glm_batch_stream",
    "b200glm_peer_export", "b200glm_peer_connect", "b200glm_lgamma_sum_local", "b200glm_set_lgamma_sum_total",
    "b200glm_glm_lpmf", "b200glm_glm_lpmf_rows", "b200glm_shard_constants_local", "b200glm_set_shard_constants_total",
    "b200glm_timeline_enable", "b200glm_timeline_read", "b200glm_abi_version", "b200glm_measure_peaks",
    "b200glm_launch_count", "b200glm_bytes_per_gradient", "b200glm_last_error", "b200glm_version",
    "b200glm_nuts_reserve", "b200glm_nuts_buffers", "b200glm_nuts_init_chain", "b200glm_nuts_round",
]


class Desc(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("K", C.c_int32), ("N", C.c_int64),
        ("X", C.c_void_p), ("ldx", C.c_int64),
        ("y_int", C.c_void_p), ("y_real", C.c_void_p),
        ("G", C.c_int32), ("data_on_device", C.c_int32), ("group", C.c_void_p),
        ("prior_alpha_sd", C.c_double), ("prior_beta_sd", C.c_double),
        ("prior_sigma_loc", C.c_double), ("prior_sigma_scale", C.c_double),
        ("prior_sigma_a_scale", C.c_double),
        ("device", C.c_int32), ("n_slots", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("N_total", C.c_int64), ("grid_ctas", C.c_int32), ("flags", C.c_int32),
        ("trials", C.c_void_p), ("n_classes", C.c_int32),
    ]


_lib = None


class NutsConfig(C.Structure):
    """b200glm_nuts_config (include/b200glm.h): device-side NUTS, SURVEY 8f row 2"""
    _fields_ = [("max_depth", C.c_int32), ("num_warmup", C.c_int32), ("num_samples", C.c_int32),
                ("w_num_warmup", C.c_uint32), ("w_init_buffer", C.c_uint32), ("w_term_buffer", C.c_uint32),
                ("w_base_window", C.c_uint32), ("w_size0", C.c_uint32), ("w_next0", C.c_uint32),
                ("max_deltaH", C.c_double), ("delta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double),
                ("t0", C.c_double), ("stepsize_jitter", C.c_double)]


class NutsStatus(C.Structure):
    """b200glm_nuts_status"""
    _fields_ = [("phase", C.c_int32), ("need_normals", C.c_int32), ("iter", C.c_int32), ("fail_code", C.c_int32),
                ("adapt_done", C.c_int32), ("reserved", C.c_int32), ("n_unif", C.c_uint64), ("eps_nom", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m stan_b200.build` "
                "(there is no CPU fallback for the GLM hot path)")
        L = C.CDLL(LIB_PATH)
        try:
            L.b200glm_abi_version.restype = C.c_int32
            abi = L.b200glm_abi_version()
        except AttributeError:
            abi = None
        if abi != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} was built for ABI {abi}, this binding needs {ABI_VERSION}: "
                               "rebuild with `python -m stan_b200.build`")
        dp = C.POINTER(C.c_double)
        L.b200glm_create.argtypes = [C.POINTER(Desc), C.POINTER(C.c_void_p)]
        L.b200glm_destroy.argtypes = [C.c_void_p]
        L.b200glm_append_rows.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        L.b200glm_finalize.argtypes = [C.c_void_p]
        L.b200glm_destroy.restype = None
        L.b200glm_num_params.argtypes = [C.c_void_p]
        L.b200glm_log_prob_grad.argtypes = [C.c_void_p, C.c_int32, dp, C.c_int32, C.c_int32, dp, dp]
        L.b200glm_log_prob.argtypes = [C.c_void_p, C.c_int32, dp, C.c_int32, C.c_int32, dp]
        L.b200glm_set_state.argtypes = [C.c_void_p, C.c_int32, dp, dp, dp, C.c_double]
        L.b200glm_leapfrog.argtypes = [C.c_void_p, C.c_int32, C.c_double, dp, dp, dp, dp, dp]
        L.b200glm_leapfrog_async.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        L.b200glm_grad_async.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.b200glm_sync.argtypes = [C.c_void_p, C.c_int32]
        L.b200glm_stream.argtypes = [C.c_void_p, C.c_int32]
        L.b200glm_stream.restype = C.c_void_p
        L.b200glm_result_device.argtypes = [C.c_void_p, C.c_int32]
        L.b200glm_result_device.restype = C.c_void_p
        L.b200glm_comm_unique_id.argtypes = [C.c_void_p]
        L.b200glm_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        ip = C.POINTER(C.c_int32)
        L.b200glm_batch_reserve.argtypes = [C.c_void_p, C.c_int32]
        L.b200glm_log_prob_grad_batched.argtypes = [C.c_void_p, C.c_int32, dp, C.c_int32, C.c_int32, dp, dp, ip]
        L.b200glm_set_state_batched.argtypes = [C.c_void_p, C.c_int32, ip, dp, dp, dp, dp, dp]
        L.b200glm_leapfrog_batched.argtypes = [C.c_void_p, C.c_int32, ip, dp, dp, dp, dp, dp, ip]
        L.b200glm_leapfrog_batched_async.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        L.b200glm_batch_sync.argtypes = [C.c_void_p]
        L.b200glm_batch_stream.argtypes = [C.c_void_p]
        L.b200glm_batch_stream.restype = C.c_void_p
        L.b200glm_peer_export.argtypes = [C.c_void_p, C.c_void_p]
        L.b200glm_peer_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.b200glm_lgamma_sum_local.argtypes = [C.c_void_p]
        L.b200glm_lgamma_sum_local.restype = C.c_double
        L.b200glm_set_lgamma_sum_total.argtypes = [C.c_void_p, C.c_double]
        L.b200glm_glm_lpmf.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, dp, C.c_double,
                                       dp, dp, dp, dp]
        L.b200glm_glm_lpmf_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, C.c_double, dp,
                                            dp, C.c_double, dp, dp, dp, dp, dp, dp]
        L.b200glm_shard_constants_local.argtypes = [C.c_void_p, dp]
        L.b200glm_set_shard_constants_total.argtypes = [C.c_void_p, dp]
        L.b200glm_timeline_enable.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.b200glm_timeline_read.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64), ip]
        L.b200glm_measure_peaks.argtypes = [C.c_int32, dp, dp]
        L.b200glm_nuts_reserve.argtypes = [C.c_void_p, C.c_int32, C.POINTER(NutsConfig)]
        L.b200glm_nuts_buffers.argtypes = [C.c_void_p, C.POINTER(dp), C.POINTER(dp), C.POINTER(C.POINTER(NutsStatus)),
                                           C.POINTER(dp), C.POINTER(dp)]
        L.b200glm_nuts_init_chain.argtypes = [C.c_void_p, C.c_int32, dp, dp, C.c_double]
        L.b200glm_nuts_round.argtypes = [C.c_void_p, C.c_int32, ip]
        L.b200glm_launch_count.argtypes = [C.c_void_p]
        L.b200glm_launch_count.restype = C.c_int64
        L.b200glm_bytes_per_gradient.argtypes = [C.c_void_p]
        L.b200glm_bytes_per_gradient.restype = C.c_int64
        L.b200glm_last_error.argtypes = [C.c_void_p]
        L.b200glm_last_error.restype = C.c_char_p
        L.b200glm_version.restype = C.c_char_p
        _lib = L
    return _lib


def connect_peers_torch(handle, world, dist, dev):
    """Wire the peer mailboxes of a row-sharded handle (b200glm_handle*) using torch.distributed as the
    out-of-band channel: all-gather the 64-byte IPC handles (rank order) and sum the per-shard create-time
    constants (propto=false constant, data-check flag).  torch.distributed is plumbing here; the exchange itself happens inside the gradient launch."""
    import torch
    L = lib()

    def check(rc):
        if rc != OK:
            raise RuntimeError((L.b200glm_last_error(handle) or b"b200glm error").decode())

    buf = C.create_string_buffer(64)
    check(L.b200glm_peer_export(handle, buf))
    mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
    check(L.b200glm_peer_connect(handle, C.create_string_buffer(raw, len(raw)), world))
    loc = (C.c_double * 2)()
    check(L.b200glm_shard_constants_local(handle, loc))
    lg = torch.tensor([loc[0], loc[1]], dtype=torch.float64, device=dev)
    dist.all_reduce(lg)
    tot = (C.c_double * 2)(float(lg[0].item()), float(lg[1].item()))
    check(L.b200glm_set_shard_constants_total(handle, tot))
    dist.barrier()


def comm_init_torch(handle, rank, world, dist, dev):
    """b200glm_comm_init for torch.distributed callers: rank 0 obtains the ncclUniqueId, torch.distributed broadcasts it
    (plumbing), every rank joins the handle's own NCCL communicator.  Needed by the NCCL transport of the single-chain
    path and by batched chains on row shards (one all-reduce of the partial sums per batched evaluation)."""
    import torch
    L = lib()
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = C.create_string_buffer(128)
        if L.b200glm_comm_unique_id(buf) == OK:      # on failure the zero id is broadcast: EVERY rank raises below
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
    dist.broadcast(uid, 0)
    raw = bytes(uid.cpu().numpy().tobytes())
    if not any(raw):
        raise RuntimeError("ncclGetUniqueId failed on rank 0 (libnccl.so.2 not loadable?)")
    if L.b200glm_comm_init(handle, C.create_string_buffer(raw, 128), rank, world) != OK:
        raise RuntimeError((L.b200glm_last_error(handle) or b"b200glm_comm_init failed").decode())


def measure_peaks(device=0, read=True, dmma=False):
    """(read-only-stream HBM GB/s, fp64 DMMA TFLOP/s) measured now on `device`; None for what was not asked."""
    r, t = C.c_double(), C.c_double()
    rc = lib().b200glm_measure_peaks(int(device), C.byref(r) if read else None, C.byref(t) if dmma else None)
    if rc != OK:
        raise RuntimeError("b200glm_measure_peaks failed (no sm_100 device?)")
    return (r.value if read else None), (t.value if dmma else None)
