"""ctypes binding of stan_b200/lib/libb200stan.so: the reference's own services
(stan::services::sample::hmc_nuts_diag_e_adapt, stan::model::log_prob_grad / gradient,
stan::mcmc::expl_leapfrog) compiled against b200::glm_model (stan_b200/cpp/b200/stan_glm_model.hpp).

This is the executable form of the drop-in claim: NUTS, adaptation, RNG, writers are the
reference's unmodified code; only the model gradient / leapfrog run on the B200.
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from .model import DomainError, InvalidArgument, CudaError, DEFAULT_PRIORS, make_desc

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb200stan.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: build with `make -C stan_b200/cpp` where the reference headers exist")
        _capi.lib()  # libb200glm.so first (same directory, also found through $ORIGIN)
        L = C.CDLL(LIB_PATH)
        L.b200stan_create.restype = C.c_void_p
        L.b200stan_create.argtypes = [C.POINTER(_capi.Desc), C.c_char_p, C.c_int]
        L.b200stan_destroy.argtypes = [C.c_void_p]
        L.b200stan_num_params.argtypes = [C.c_void_p]
        L.b200stan_backend_handle.restype = C.c_void_p
        L.b200stan_backend_handle.argtypes = [C.c_void_p]
        L.b200stan_version.restype = C.c_char_p
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class StanGLM:
    """b200::glm_model driven through the reference's C++ interfaces."""

    def __init__(self, family, X, y, group=None, G=0, device=0, n_slots=8, rank=0, world=1, N_total=0,
                 data_on_device=False, N=None, K=None, ldx=None, n_classes=0, **priors):
        """Arguments as GLMModel.  rank/world/N_total: this process holds row shard `rank` of `world` (one
        process per GPU); call connect_peers_torch() before the first evaluation and use n_slots=1 and one
        host thread: every rank must issue the same sequence of evaluations, which the reference's
        deterministic host code does by construction when it is given the same seeds."""
        self.L = lib()
        d, keep = make_desc(family, X, y, group, G, device, n_slots, rank, world, N_total,
                            data_on_device=data_on_device, N=N, K=K, ldx=ldx, n_classes=n_classes, **priors)
        self.rank, self.world = int(rank), int(world)
        err = C.create_string_buffer(1024)
        self.h = C.c_void_p(self.L.b200stan_create(C.byref(d), err, 1024))
        if not self.h:
            raise CudaError(err.value.decode() or "b200stan_create failed")
        self.P = self.L.b200stan_num_params(self.h)

    @classmethod
    def from_json(cls, path, family, name_y="y", name_X="X", center_x=False, device=0, n_slots=8, **priors):
        """The stanc-style constructor b200::glm_model(var_context&, glm_config): the data block is parsed by
        the reference's own stan::json::json_data from a CmdStan-format JSON file."""
        self = cls.__new__(cls)
        self.L = lib()
        self.rank, self.world = 0, 1
        pri = dict(DEFAULT_PRIORS)
        pri.update(priors)
        f = self.L.b200stan_create_from_json
        f.restype = C.c_void_p
        f.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int] + [C.c_double] * 5 + [C.c_int, C.c_int,
                                                                                                   C.c_char_p, C.c_int]
        err = C.create_string_buffer(1024)
        self.h = C.c_void_p(f(os.fsencode(path), _capi.FAMILY[family], name_y.encode(), name_X.encode(), int(center_x),
                              pri["prior_alpha_sd"], pri["prior_beta_sd"], pri["prior_sigma_loc"],
                              pri["prior_sigma_scale"], pri["prior_sigma_a_scale"], device, n_slots, err, 1024))
        if not self.h:
            raise InvalidArgument(err.value.decode() or "b200stan_create_from_json failed")
        self.P = self.L.b200stan_num_params(self.h)
        return self

    @classmethod
    def from_dump(cls, path, family, device=0, n_slots=8):
        """The same constructor fed by the reference's R-dump reader stan::io::dump."""
        self = cls.__new__(cls)
        self.L = lib()
        self.rank, self.world = 0, 1
        f = self.L.b200stan_create_from_dump
        f.restype = C.c_void_p
        f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        err = C.create_string_buffer(1024)
        self.h = C.c_void_p(f(os.fsencode(path), _capi.FAMILY[family], device, n_slots, err, 1024))
        if not self.h:
            raise InvalidArgument(err.value.decode() or "b200stan_create_from_dump failed")
        self.P = self.L.b200stan_num_params(self.h)
        return self

    def log_prob_propto(self, theta, jacobian=True, eigen=False):
        """stan::model::log_prob_propto<jacobian>(model, theta) -- both signatures of log_prob_propto.hpp."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp, err = C.c_double(), C.create_string_buffer(1024)
        rc = self.L.b200stan_log_prob_propto(self.h, _dp(th), int(jacobian), int(eigen), C.byref(lp), err, 1024)
        if rc:
            self._raise(rc, err)
        return lp.value

    def means_x(self):
        """Column means removed from X by center_x (brms-style transformed data); empty if not centred."""
        self.L.b200stan_means_x.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        out = np.empty(max(self.P, 1))
        n = self.L.b200stan_means_x(self.h, _dp(out))
        return out[:n].copy()

    def connect_peers_torch(self, dist, dev):
        """Row-sharded model: map every rank's mailbox (the likelihood partials are then exchanged inside
        the gradient / leapfrog launch over NVLink; no collective call per gradient)."""
        _capi.connect_peers_torch(C.c_void_p(self.L.b200stan_backend_handle(self.h)), self.world, dist, dev)

    def comm_init_torch(self, dist, dev):
        """Row-sharded model: join the backend handle's NCCL communicator (needed by nuts_device / nuts_batched on row
        shards: the partial sums of all chains are combined by one all-reduce per round; the single-chain path keeps
        the in-kernel mailbox exchange when connect_peers_torch was called as well)."""
        _capi.comm_init_torch(C.c_void_p(self.L.b200stan_backend_handle(self.h)), self.rank, self.world, dist, dev)

    def close(self):
        if getattr(self, "h", None):
            self.L.b200stan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _raise(rc, err):
        msg = err.value.decode()
        if rc == 1:
            raise DomainError(msg)
        if rc == 2:
            raise InvalidArgument(msg)
        raise CudaError(msg)

    def counters(self):
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        self.L.b200stan_counters(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(gradients=a.value, leapfrogs=b.value, uploads=c.value)

    def log_prob_grad(self, theta, propto=True, jacobian=True):
        """stan::model::log_prob_grad<propto,jacobian> (AD tape + precomputed_gradients)."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp, g, err = C.c_double(), np.empty(self.P), C.create_string_buffer(1024)
        rc = self.L.b200stan_log_prob_grad(self.h, _dp(th), int(propto), int(jacobian), C.byref(lp), _dp(g), err, 1024)
        if rc:
            self._raise(rc, err)
        return lp.value, g

    def log_prob(self, theta, propto=False, jacobian=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp, err = C.c_double(), C.create_string_buffer(1024)
        rc = self.L.b200stan_log_prob(self.h, _dp(th), int(propto), int(jacobian), C.byref(lp), err, 1024)
        if rc:
            self._raise(rc, err)
        return lp.value

    def gradient(self, theta):
        """stan::model::gradient (the explicit specialisation: no tape)."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp, g, err = C.c_double(), np.empty(self.P), C.create_string_buffer(1024)
        rc = self.L.b200stan_gradient(self.h, _dp(th), C.byref(lp), _dp(g), err, 1024)
        if rc:
            self._raise(rc, err)
        return lp.value, g

    def leapfrog(self, eps, inv_metric, q, p, n_steps=1):
        """hamiltonian.init(z) then n_steps x expl_leapfrog<diag_e_metric<b200::glm_model>>::evolve."""
        q, p = (np.array(a, dtype=np.float64) for a in (q, p))
        g = np.zeros(self.P)
        im = np.ascontiguousarray(inv_metric, dtype=np.float64)
        V, err = C.c_double(), C.create_string_buffer(1024)
        rc = self.L.b200stan_leapfrog(self.h, C.c_double(eps), _dp(im), int(n_steps), _dp(q), _dp(p), _dp(g),
                                      C.byref(V), err, 1024)
        if rc:
            self._raise(rc, err)
        return q, p, g, V.value

    def _set_jitter(self, j):
        """the services' stepsize_jitter argument for the next nuts* call (base_hmc::set_stepsize_jitter keeps (0, 1) only)"""
        self.L.b200stan_set_stepsize_jitter.argtypes = [C.c_double]
        self.L.b200stan_set_stepsize_jitter.restype = None
        self.L.b200stan_set_stepsize_jitter(float(j))

    def nuts(self, num_chains=4, seed=1, init_chain_id=1, init_radius=2.0, num_warmup=1000, num_samples=1000,
             stepsize=1.0, max_depth=10, delta=0.8, num_threads=0, stepsize_jitter=0.0):
        """stan::services::sample::hmc_nuts_diag_e_adapt, unmodified, on b200::glm_model."""
        self._set_jitter(stepsize_jitter)
        W = 7 + self.P
        draws = np.empty((num_chains, num_warmup + num_samples, W))
        step, inv_metric = np.empty(num_chains), np.empty((num_chains, self.P))
        warm_lf, wall, err = np.empty(num_chains), C.c_double(), C.create_string_buffer(4096)
        rc = self.L.b200stan_nuts(self.h, num_chains, C.c_uint(seed), C.c_uint(init_chain_id), C.c_double(init_radius),
                                  num_warmup, num_samples, C.c_double(stepsize), max_depth, C.c_double(delta),
                                  num_threads, _dp(draws), _dp(step), _dp(inv_metric), _dp(warm_lf), C.byref(wall),
                                  err, 4096)
        if rc:
            self._raise(rc, err)
        return dict(draws=draws[:, num_warmup:, :], warmup_draws=draws[:, :num_warmup, :], stepsize=step,
                    inv_metric=inv_metric, warm_leapfrogs=warm_lf, wall=wall.value)

    def nuts_csv(self, prefix, num_chains=4, seed=1, init_chain_id=1, init_radius=2.0, num_warmup=1000,
                 num_samples=1000, stepsize=1.0, max_depth=10, delta=0.8, num_threads=0):
        """The same service writing through the reference's CmdStan-format writers: `<prefix>_<chain>.csv`
        (unique_stream_writer) and `<prefix>_metric_<chain>.json` (json_writer).  Returns the csv paths."""
        err = C.create_string_buffer(4096)
        rc = self.L.b200stan_nuts_csv(self.h, num_chains, C.c_uint(seed), C.c_uint(init_chain_id),
                                      C.c_double(init_radius), num_warmup, num_samples, C.c_double(stepsize), max_depth,
                                      C.c_double(delta), num_threads, os.fsencode(prefix), err, 4096)
        if rc:
            self._raise(rc, err)
        return [f"{prefix}_{init_chain_id + c}.csv" for c in range(num_chains)]

    def nuts_batched(self, num_chains=4, seed=1, init_chain_id=1, init_radius=2.0, num_warmup=1000, num_samples=1000,
                     stepsize=1.0, max_depth=10, delta=0.8, stepsize_jitter=0.0):
        """b200::hmc_nuts_diag_e_adapt_batched: one host thread per chain running the reference's single-chain
        service; all chains' leapfrog steps served together by one batched fp64 DMMA launch."""
        self._set_jitter(stepsize_jitter)
        W = 7 + self.P
        draws = np.empty((num_chains, num_warmup + num_samples, W))
        step, inv_metric = np.empty(num_chains), np.empty((num_chains, self.P))
        warm_lf, wall, err = np.empty(num_chains), C.c_double(), C.create_string_buffer(4096)
        stats = (C.c_long * 10)()
        rc = self.L.b200stan_nuts_batched(self.h, num_chains, C.c_uint(seed), C.c_uint(init_chain_id),
                                          C.c_double(init_radius), num_warmup, num_samples, C.c_double(stepsize),
                                          max_depth, C.c_double(delta), _dp(draws), _dp(step), _dp(inv_metric),
                                          _dp(warm_lf), C.byref(wall), stats, err, 4096)
        if rc:
            self._raise(rc, err)
        return dict(draws=draws[:, num_warmup:, :], warmup_draws=draws[:, :num_warmup, :], stepsize=step,
                    inv_metric=inv_metric, warm_leapfrogs=warm_lf, wall=wall.value,
                    batches=int(stats[0]), lanes=int(stats[1]),
                    batch_size_hist={k: int(stats[2 + i]) for i, k in enumerate(
                        ["<=2", "<=16", "<=32", "<=64", "<=128", "<=256", "<=512", ">512"])})


    def nuts_device(self, num_chains=4, seed=1, init_chain_id=1, init_radius=2.0, num_warmup=1000, num_samples=1000,
                    stepsize=1.0, max_depth=10, delta=0.8, stepsize_jitter=0.0):
        """b200::hmc_nuts_diag_e_adapt_device: the NUTS transition (iterative build_tree, U-turn checks, multinomial
        sampling) and the adaptation (dual averaging, Welford metric windows, init_stepsize) run per chain on the
        device right behind the batched leapfrog; the host keeps the chains' boost engines and collects the draws."""
        self._set_jitter(stepsize_jitter)
        W = 7 + self.P
        draws = np.empty((num_chains, num_warmup + num_samples, W))
        step, inv_metric = np.empty(num_chains), np.empty((num_chains, self.P))
        warm_lf, wall, err = np.empty(num_chains), C.c_double(), C.create_string_buffer(4096)
        stats = (C.c_long * 10)()
        rc = self.L.b200stan_nuts_device(self.h, num_chains, C.c_uint(seed), C.c_uint(init_chain_id),
                                         C.c_double(init_radius), num_warmup, num_samples, C.c_double(stepsize),
                                         max_depth, C.c_double(delta), _dp(draws), _dp(step), _dp(inv_metric),
                                         _dp(warm_lf), C.byref(wall), stats, err, 4096)
        if rc:
            self._raise(rc, err)
        return dict(draws=draws[:, num_warmup:, :], warmup_draws=draws[:, :num_warmup, :], stepsize=step,
                    inv_metric=inv_metric, warm_leapfrogs=warm_lf, wall=wall.value, rounds=int(stats[0]),
                    lanes=int(stats[1]), uniforms=int(stats[2]), normal_vectors=int(stats[3]))


class FuncGLM:
    """b200::glm_data + the stan::math overloads of b200/glm_functions.hpp (function-level binding),
    exercised as one node of a reverse-mode tape: f = scale * glm_lpmf(...) + 0.5 * sum(beta^2)."""

    def __init__(self, family, X, y, group=None, G=0, trials=None):
        self.L = lib()
        self.L.b200stan_func_create.restype = C.c_void_p
        self.L.b200stan_func_destroy.argtypes = [C.c_void_p]
        self.fam = _capi.FAMILY[family]
        X = np.asfortranarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64 if self.fam == 2 else np.int32)
        grp = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        nt = None if trials is None else np.ascontiguousarray(trials, dtype=np.int32)
        self.K, self.G = X.shape[1], int(G)
        err = C.create_string_buffer(1024)
        self.h = C.c_void_p(self.L.b200stan_func_create(
            C.c_int(self.fam), C.c_longlong(X.shape[0]), C.c_int(self.K), _dp(X), C.c_void_p(y.ctypes.data),
            None if grp is None else grp.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(self.G), err, 1024,
            None if nt is None else nt.ctypes.data_as(C.POINTER(C.c_int))))
        if not self.h:
            raise CudaError(err.value.decode() or "b200stan_func_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.L.b200stan_func_destroy(self.h)
            self.h = None

    def eval(self, alpha, beta, sigma=1.0, propto=True, operands_are_var=True, sigma_is_var=True, scale=1.0):
        a = np.ascontiguousarray(np.atleast_1d(alpha), dtype=np.float64)
        b = np.ascontiguousarray(beta, dtype=np.float64)
        f, ds, err = C.c_double(), C.c_double(), C.create_string_buffer(1024)
        da, db = np.zeros_like(a), np.zeros(max(self.K, 1))
        rc = self.L.b200stan_func_eval(self.h, int(propto), int(operands_are_var), int(sigma_is_var), _dp(a),
                                       C.c_int(a.size), _dp(b), C.c_double(float(sigma)), C.c_double(float(scale)),
                                       C.byref(f), _dp(da), _dp(db), C.byref(ds), err, 1024)
        if rc:
            StanGLM._raise(rc, err)
        return f.value, da, db[:self.K], ds.value


def func_eval_rows(fg, alpha_rows, beta, sigma_rows=None, sigma=1.0, propto=True, scale=1.0):
    """FuncGLM with the per-row overloads (b200::by_row(alpha), vector sigma), every operand a var:
    returns f = scale * glm(...) + 0.5 * sum(beta^2) and its adjoints (d_alpha_rows, d_beta, d_sigma_rows | d_sigma)."""
    L = fg.L
    a = np.ascontiguousarray(alpha_rows, dtype=np.float64)
    b = np.ascontiguousarray(beta, dtype=np.float64)
    s = None if sigma_rows is None else np.ascontiguousarray(sigma_rows, dtype=np.float64)
    f, ds, err = C.c_double(), C.c_double(), C.create_string_buffer(1024)
    da, db = np.zeros_like(a), np.zeros(max(fg.K, 1))
    dsr = np.zeros_like(a)
    rc = L.b200stan_func_eval_rows(fg.h, int(propto), _dp(a), _dp(b), None if s is None else _dp(s),
                                   C.c_double(float(sigma)), C.c_double(float(scale)), C.byref(f), _dp(da), _dp(db),
                                   _dp(dsr), C.byref(ds), err, 1024)
    if rc:
        StanGLM._raise(rc, err)
    return f.value, da, db[:fg.K], (dsr if s is not None else ds.value)


def diagnostic(which, draws):
    """stan::analyze::{ess, rhat, mcse_mean, mcse_sd} (the reference's estimators, compiled into libb200stan.so)
    for ONE parameter; draws: (n_draws, n_chains)."""
    L = lib()
    L.b200stan_diagnostic.restype = C.c_double
    L.b200stan_diagnostic.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int]
    d = np.asfortranarray(draws, dtype=np.float64)
    return L.b200stan_diagnostic({"ess": 0, "rhat": 1, "mcse_mean": 2, "mcse_sd": 3}[which], _dp(d), d.shape[0],
                                 d.shape[1])


def read_stan_csv(path):
    """stan::io::stan_csv_reader::parse: returns dict(header, samples (rows, cols), step_size, metric)."""
    L = lib()
    n_rows, n_cols, n_metric, step = C.c_int(), C.c_int(), C.c_int(), C.c_double()
    header, err = C.create_string_buffer(1 << 20), C.create_string_buffer(1024)
    metric = np.empty(1 << 16)
    args = lambda buf, mr, mc: (os.fsencode(path), buf, mr, mc, C.byref(n_rows), C.byref(n_cols), C.byref(step),
                                _dp(metric), metric.size, C.byref(n_metric), header, len(header), err, 1024)
    rc = L.b200stan_read_csv(*args(None, 0, 0))
    if rc:
        StanGLM._raise(rc, err)
    samples = np.empty((n_rows.value, n_cols.value))
    rc = L.b200stan_read_csv(*args(_dp(samples) if samples.size else None, n_rows.value, n_cols.value))
    if rc:
        StanGLM._raise(rc, err)
    return dict(header=header.value.decode().split(","), samples=samples, step_size=step.value,
                metric=metric[:n_metric.value].copy())
