"""Host-side mirror of the reference's model interface for the GLM hot path.

`GLMModel` plays the role of a stanc-generated model class (stan::model::model_base_crtp,
src/stan/model/model_base_crtp.hpp:74-301) whose log_prob is evaluated on the B200:
  num_params_r()                    prob_grad.hpp:19-84
  log_prob_grad(theta, propto, jac) stan::model::log_prob_grad  (log_prob_grad.hpp:29-50)
  log_prob(theta, propto, jac)      Model::log_prob<propto,jacobian>(double)  (initialize.hpp:128)
  leapfrog(...) / set_state(...)    expl_leapfrog::evolve on device-resident z (base_leapfrog.hpp:17-22)
Error behaviour follows the reference: DomainError <-> std::domain_error (recoverable),
InvalidArgument <-> std::invalid_argument, CudaError <-> any other std::exception (fatal).
Everything goes through the C ABI in include/b200glm.h; there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi


class DomainError(ValueError):
    """std::domain_error of the reference (proposal rejected / init retried)."""


class InvalidArgument(ValueError):
    """std::invalid_argument of the reference (size mismatch etc.)."""


class CudaError(RuntimeError):
    """CUDA / NCCL failure or missing device: fatal, never falls back to a CPU path."""


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


DEFAULT_PRIORS = dict(prior_alpha_sd=2.5, prior_beta_sd=2.5, prior_sigma_loc=1.0,
                      prior_sigma_scale=2.0, prior_sigma_a_scale=1.0)


def make_desc(family, X, y, group=None, G=0, device=0, n_slots=1, rank=0, world=1, N_total=0,
              grid_ctas=0, flags=0, data_on_device=False, N=None, K=None, ldx=None, trials=None, n_classes=0, **priors):
    """Fill a b200glm_desc.  X, y, group: numpy arrays (host), or -- with data_on_device=True -- integer
    device pointers (e.g. torch tensor .data_ptr()) with N, K, ldx given explicitly.  Returns (desc, keep):
    `keep` holds the host arrays the descriptor points into (they must outlive the create call)."""
    fam = _capi.FAMILY[family]
    d = _capi.Desc()
    keep = []
    if data_on_device:
        d.N, d.K, d.ldx = int(N), int(K), int(ldx if ldx is not None else N)
        d.X = int(X) if X else None
        if fam == 2:
            d.y_real, d.y_int = (int(y) if y else None), None
        else:
            d.y_int, d.y_real = (int(y) if y else None), None
        d.group = int(group) if G else None
        d.trials = int(trials) if (fam == 3 and trials) else None
    else:
        X = np.asfortranarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise InvalidArgument("X must be a matrix")
        keep.append(X)
        d.N, d.K = X.shape
        d.ldx = max(X.shape[0], 1)
        d.X = X.ctypes.data if X.size else None
        y = np.ascontiguousarray(y, dtype=np.float64 if fam == 2 else np.int32)
        if y.shape != (d.N,):
            raise InvalidArgument("Vector of dependent variables has the wrong size")
        keep.append(y)
        if fam == 2:
            d.y_real, d.y_int = (y.ctypes.data if y.size else None), None
        else:
            d.y_int, d.y_real = (y.ctypes.data if y.size else None), None
        if G:
            group = np.ascontiguousarray(group, dtype=np.int32)
            if group.shape != (d.N,):
                raise InvalidArgument("Vector of intercepts has the wrong size")
            keep.append(group)
            d.group = group.ctypes.data if group.size else None
        if fam == 3:     # binomial_logit: population sizes (binomial_logit_glm_lpmf.hpp:93-97 check_consistent_sizes)
            trials = np.ascontiguousarray(trials, dtype=np.int32)
            if trials.shape != (d.N,):
                raise InvalidArgument("Population size parameter has the wrong size")
            keep.append(trials)
            d.trials = trials.ctypes.data if trials.size else None
    d.family, d.G, d.data_on_device = fam, int(G), int(bool(data_on_device))
    d.n_classes = int(n_classes)
    pri = dict(DEFAULT_PRIORS)
    pri.update(priors)
    for k, v in pri.items():
        setattr(d, k, float(v))
    d.device, d.n_slots, d.rank, d.world = int(device), int(n_slots), int(rank), int(world)
    d.N_total, d.grid_ctas, d.flags = int(N_total), int(grid_ctas), int(flags)
    return d, keep


class GLMModel:
    def __init__(self, family, X, y, group=None, G=0, device=0, n_slots=1, rank=0, world=1, N_total=0,
                 grid_ctas=0, flags=0, data_on_device=False, N=None, K=None, ldx=None, trials=None,
                 n_classes=0, _host_chunks=False, **priors):
        """X, y, group, trials: numpy arrays (host), or -- with data_on_device=True -- integer device
        pointers (e.g. torch tensor .data_ptr()) with N, K, ldx given explicitly."""
        self.L = _capi.lib()
        self.family = family
        d, keep = make_desc(family, X, y, group, G, device, n_slots, rank, world, N_total, grid_ctas, flags,
                            data_on_device, N, K, ldx, trials, n_classes, **priors)
        if _host_chunks:
            d.data_on_device = 0
        self.N, self.K, self.G, self.n_classes = int(d.N), int(d.K), int(G), int(n_classes)
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        rc = self.L.b200glm_create(C.byref(d), C.byref(h))
        self.h = h
        try:
            self._check(rc)
        except Exception:
            self.close()       # a failed create still returns a handle (for last_error); release it
            raise
        self.P = self.L.b200glm_num_params(self.h)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc == _capi.OK:
            return
        msg = self.L.b200glm_last_error(self.h).decode() if self.h else "b200glm_create failed"
        if rc == _capi.DOMAIN:
            raise DomainError(msg)
        if rc == _capi.INVALID:
            raise InvalidArgument(msg)
        raise CudaError(msg)

    # ------------------------------------------------------------------ streamed construction
    @classmethod
    def streamed(cls, family, N, K, device=0, data_on_device=True, **kw):
        """A handle with room for N rows and no data yet (B200GLM_FLAG_STREAMED): feed it with append_rows(), then
        finalize().  X is then never resident a second time next to its panel copy."""
        return cls(family, None, None, device=device, data_on_device=True, N=N, K=K, ldx=max(N, 1),
                   flags=kw.pop("flags", 0) | 2, _host_chunks=not data_on_device, **kw)

    def append_rows(self, X, y, trials=None, n=None, ldx=None):
        """Next rows: device pointers (ints) with n / ldx given, or numpy arrays (column-major X) for a handle created
        with data_on_device=False."""
        if isinstance(X, (int, type(None))) and not isinstance(y, np.ndarray):
            Xp, yp, tp = X, y, trials
        else:
            X = np.asfortranarray(X, dtype=np.float64)
            y = np.ascontiguousarray(y, dtype=np.float64 if self.family == "normal_id" else np.int32)
            n, ldx = X.shape[0], max(X.shape[0], 1)
            Xp, yp = X.ctypes.data, y.ctypes.data
            tp = None
            if trials is not None:
                trials = np.ascontiguousarray(trials, dtype=np.int32)
                tp = trials.ctypes.data
        real = self.family == "normal_id"
        self._check(self.L.b200glm_append_rows(self.h, int(n), Xp, int(ldx if ldx is not None else n),
                                               None if real else yp, yp if real else None, tp))

    def finalize(self):
        self._check(self.L.b200glm_finalize(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.b200glm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ model API
    def num_params_r(self):
        return self.P

    def param_names(self):
        if self.family == "ordered_logistic":
            return [f"beta.{k}" for k in range(1, self.K + 1)] + [f"c.{j}" for j in range(1, self.n_classes)]
        if self.family == "categorical_logit":
            return ([f"alpha.{c}" for c in range(1, self.n_classes + 1)]
                    + [f"beta.{k}.{c}" for c in range(1, self.n_classes + 1) for k in range(1, self.K + 1)])
        n = (["mu_a", "sigma_a"] + [f"a.{g}" for g in range(1, self.G + 1)]) if self.G else ["alpha"]
        n += [f"beta.{k}" for k in range(1, self.K + 1)]
        if self.family == "normal_id":
            n.append("sigma")
        if self.family == "neg_binomial_2_log":
            n.append("phi")
        return n

    def _theta(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        if th.shape != (self.P,):
            raise InvalidArgument(f"theta has {th.size} entries, model has {self.P} parameters")
        return th

    def log_prob_grad(self, theta, propto=True, jacobian=True, slot=0):
        th = self._theta(theta)
        lp = C.c_double()
        g = np.empty(self.P)
        self._check(self.L.b200glm_log_prob_grad(self.h, slot, _dp(th), int(propto), int(jacobian),
                                                 C.byref(lp), _dp(g)))
        return lp.value, g

    def log_prob(self, theta, propto=False, jacobian=True, slot=0):
        th = self._theta(theta)
        lp = C.c_double()
        self._check(self.L.b200glm_log_prob(self.h, slot, _dp(th), int(propto), int(jacobian), C.byref(lp)))
        return lp.value

    def glm_lpmf(self, alpha, beta, sigma=1.0, propto=True, operands_are_var=True, sigma_is_var=True, slot=0):
        """Function-level entry: the GLM density term alone and its partials (d_alpha, d_beta, d_sigma)."""
        a = np.ascontiguousarray(np.atleast_1d(alpha), dtype=np.float64)
        b = np.ascontiguousarray(beta, dtype=np.float64)
        if a.shape != (max(self.G, 1),) or b.shape != (self.K,):
            raise InvalidArgument("alpha / beta have the wrong size")
        lp, ds = C.c_double(), C.c_double()
        da, db = np.empty_like(a), np.empty(max(self.K, 1))
        self._check(self.L.b200glm_glm_lpmf(self.h, slot, int(propto), int(operands_are_var), int(sigma_is_var),
                                            _dp(a), _dp(b), float(sigma), C.byref(lp), _dp(da), _dp(db), C.byref(ds)))
        return lp.value, da, db[:self.K], ds.value

    def glm_lpmf_rows(self, beta, alpha=0.0, sigma=1.0, alpha_rows=None, sigma_rows=None, propto=True,
                      operands_are_var=True, sigma_is_var=True, slot=0):
        """Function-level entry with per-row operands: an N-vector intercept (alpha_rows) and / or, for normal_id, an
        N-vector scale (sigma_rows).  Returns lp, d_alpha (scalar or N-vector), d_beta, d_sigma (scalar or N-vector)."""
        b = np.ascontiguousarray(beta, dtype=np.float64)
        if b.shape != (self.K,):
            raise InvalidArgument("beta has the wrong size")
        ar = None if alpha_rows is None else np.ascontiguousarray(alpha_rows, dtype=np.float64)
        sr = None if sigma_rows is None else np.ascontiguousarray(sigma_rows, dtype=np.float64)
        for v, what in ((ar, "Vector of intercepts"), (sr, "Scale vector")):
            if v is not None and v.shape != (self.N,):
                raise InvalidArgument(f"{what} has the wrong size")       # check_size_match(N, size(alpha | sigma))
        lp, da, ds = C.c_double(), C.c_double(), C.c_double()
        db = np.empty(max(self.K, 1))
        dar = None if ar is None else np.empty(self.N)
        dsr = None if sr is None else np.empty(self.N)
        self._check(self.L.b200glm_glm_lpmf_rows(
            self.h, slot, int(propto), int(operands_are_var), int(sigma_is_var), None if ar is None else _dp(ar),
            float(alpha), _dp(b), None if sr is None else _dp(sr), float(sigma), C.byref(lp),
            None if dar is None else _dp(dar), C.byref(da), _dp(db), None if dsr is None else _dp(dsr), C.byref(ds)))
        return lp.value, (da.value if dar is None else dar), db[:self.K], (ds.value if dsr is None else dsr)

    def set_state(self, q, p, g, V, slot=0):
        q, p, g = (self._theta(a) for a in (q, p, g))
        self._check(self.L.b200glm_set_state(self.h, slot, _dp(q), _dp(p), _dp(g), float(V)))

    def leapfrog(self, eps, inv_metric=None, slot=0):
        q, p, g = np.empty(self.P), np.empty(self.P), np.empty(self.P)
        V = C.c_double()
        im = None if inv_metric is None else self._theta(inv_metric)
        self._check(self.L.b200glm_leapfrog(self.h, slot, float(eps), None if im is None else _dp(im),
                                            _dp(q), _dp(p), _dp(g), C.byref(V)))
        return q, p, g, V.value

    # ------------------------------------------------------------------ async / device-resident
    def leapfrog_async(self, eps, slot=0):
        self._check(self.L.b200glm_leapfrog_async(self.h, slot, float(eps)))

    def grad_async(self, theta_device_ptr=None, slot=0):
        self._check(self.L.b200glm_grad_async(self.h, slot, theta_device_ptr))

    def sync(self, slot=0):
        self._check(self.L.b200glm_sync(self.h, slot))

    def stream_ptr(self, slot=0):
        return self.L.b200glm_stream(self.h, slot)

    def launch_count(self):
        return self.L.b200glm_launch_count(self.h)

    def bytes_per_gradient(self):
        return self.L.b200glm_bytes_per_gradient(self.h)

    def timeline_enable(self, on=True, slot=0):
        self._check(self.L.b200glm_timeline_enable(self.h, slot, int(on)))

    def timeline_read(self, slot=0):
        """Per-phase time stamps of the slot's last gradient launch: uint64 array (grid + 1, 16), see b200glm.h."""
        rows = C.c_int32()
        self._check(self.L.b200glm_timeline_read(self.h, slot, None, C.byref(rows)))
        out = np.zeros((rows.value, 16), dtype=np.uint64)
        self._check(self.L.b200glm_timeline_read(self.h, slot, out.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                 C.byref(rows)))
        return out

    # ------------------------------------------------------------------ batched chains (fp64 DMMA path)
    def batch_reserve(self, max_chains):
        self._check(self.L.b200glm_batch_reserve(self.h, int(max_chains)))
        self.max_chains = int(max_chains)

    def _mat(self, a, n):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (n, self.P):
            raise InvalidArgument(f"expected a ({n}, {self.P}) chain-major array, got {a.shape}")
        return a

    @staticmethod
    def _ip(a):
        return a.ctypes.data_as(C.POINTER(C.c_int32))

    def log_prob_grad_batched(self, thetas, propto=True, jacobian=True):
        """n chains in one pass over X: returns lp (n,), grad (n, P), status (n,) int32."""
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        n = th.shape[0]
        th = self._mat(th, n)
        lp, g, st = np.empty(n), np.empty((n, self.P)), np.empty(n, dtype=np.int32)
        self._check(self.L.b200glm_log_prob_grad_batched(self.h, n, _dp(th), int(propto), int(jacobian),
                                                         _dp(lp), _dp(g), self._ip(st)))
        return lp, g, st

    def set_state_batched(self, q, p, g, V, inv_metric=None, chains=None):
        n = np.asarray(q).shape[0]
        q, p, g = self._mat(q, n), self._mat(p, n), self._mat(g, n)
        V = np.ascontiguousarray(V, dtype=np.float64)
        im = None if inv_metric is None else self._mat(inv_metric, n)
        ch = None if chains is None else np.ascontiguousarray(chains, dtype=np.int32)
        self._check(self.L.b200glm_set_state_batched(self.h, n, None if ch is None else self._ip(ch), _dp(q), _dp(p),
                                                     _dp(g), _dp(V), None if im is None else _dp(im)))

    def leapfrog_batched(self, eps, chains=None):
        eps = np.ascontiguousarray(eps, dtype=np.float64)
        n = eps.shape[0]
        ch = None if chains is None else np.ascontiguousarray(chains, dtype=np.int32)
        q, p, g = (np.empty((n, self.P)) for _ in range(3))
        V, st = np.empty(n), np.empty(n, dtype=np.int32)
        self._check(self.L.b200glm_leapfrog_batched(self.h, n, None if ch is None else self._ip(ch), _dp(eps),
                                                    _dp(q), _dp(p), _dp(g), _dp(V), self._ip(st)))
        return q, p, g, V, st

    def leapfrog_batched_async(self, n, eps):
        self._check(self.L.b200glm_leapfrog_batched_async(self.h, int(n), float(eps)))

    def batch_sync(self):
        self._check(self.L.b200glm_batch_sync(self.h))

    def batch_stream_ptr(self):
        return self.L.b200glm_batch_stream(self.h)

    # ------------------------------------------------------------------ multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        if _capi.lib().b200glm_comm_unique_id(buf) != 0:
            raise CudaError("ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.L.b200glm_comm_init(self.h, buf, self.rank, self.world))

    def comm_init_torch(self, dist, dev):
        """comm_init with the unique id broadcast over torch.distributed (plumbing)."""
        _capi.comm_init_torch(self.h, self.rank, self.world, dist, dev)

    def peer_export(self):
        """64-byte IPC handle of this rank's mailbox (gather over ranks, then peer_connect)."""
        buf = C.create_string_buffer(64)
        self._check(self.L.b200glm_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, all_handles):
        """all_handles: world x 64 bytes in rank order."""
        raw = bytes(all_handles)
        if len(raw) != 64 * self.world:
            raise InvalidArgument("expected world x 64 bytes of IPC handles")
        self._check(self.L.b200glm_peer_connect(self.h, C.create_string_buffer(raw, len(raw)), self.world))

    def lgamma_sum_local(self):
        return self.L.b200glm_lgamma_sum_local(self.h)

    def set_lgamma_sum_total(self, total):
        self._check(self.L.b200glm_set_lgamma_sum_total(self.h, float(total)))

    def connect_peers_torch(self, dist, dev):
        """Convenience for torch.distributed callers: all-gather the mailbox handles (and the poisson
        constant) and connect.  torch.distributed is plumbing here, the exchange itself is in-kernel."""
        _capi.connect_peers_torch(self.h, self.world, dist, dev)
