"""ctypes loaders for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/glm_oracle.h).

  PortOracle : oracle/_build/libglm_oracle.so   (plain-C restatement, oracle/glm_oracle.c)
  RefOracle  : oracle/_ref/libref_oracle_v{4,3}.so (the reference itself, compiled from
               /root/reference by oracle/Makefile; absent => RefOracle.available() is False)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BERNOULLI_LOGIT, POISSON_LOG, NORMAL_ID, BINOMIAL_LOGIT, NEG_BINOMIAL_2_LOG = 0, 1, 2, 3, 4
ORDERED_LOGISTIC, CATEGORICAL_LOGIT = 5, 6       # oracle only so far: not built on the device
FAMILY = {"bernoulli_logit": 0, "poisson_log": 1, "normal_id": 2, "binomial_logit": 3, "neg_binomial_2_log": 4,
          "ordered_logistic": 5, "categorical_logit": 6}
HAS_SCALE = (NORMAL_ID, NEG_BINOMIAL_2_LOG)      # families with a trailing positive scalar (sigma | phi)


class GlmSpec(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("K", C.c_int32), ("N", C.c_int64),
        ("X", C.c_void_p), ("ldx", C.c_int64),
        ("y_int", C.c_void_p), ("y_real", C.c_void_p),
        ("G", C.c_int32), ("_pad", C.c_int32), ("group", C.c_void_p),
        ("prior_alpha_sd", C.c_double), ("prior_beta_sd", C.c_double),
        ("prior_sigma_loc", C.c_double), ("prior_sigma_scale", C.c_double),
        ("prior_sigma_a_scale", C.c_double), ("trials", C.c_void_p),
        ("n_classes", C.c_int32), ("_pad2", C.c_int32),
    ]


DEFAULT_PRIORS = dict(prior_alpha_sd=2.5, prior_beta_sd=2.5, prior_sigma_loc=1.0,
                      prior_sigma_scale=2.0, prior_sigma_a_scale=1.0)


def make_spec(family, X, y, group=None, G=0, trials=None, n_classes=0, **priors):
    """Returns (spec, keepalive).  X: (N,K) float64, any layout (copied to Fortran order)."""
    fam = FAMILY[family] if isinstance(family, str) else int(family)
    X = np.asfortranarray(X, dtype=np.float64)
    N, K = X.shape
    keep = [X]
    s = GlmSpec()
    s.family, s.K, s.N = fam, K, N
    s.X = X.ctypes.data if X.size else None
    s.ldx = max(N, 1) if X.size == 0 else N
    if fam == NORMAL_ID:
        yr = np.ascontiguousarray(y, dtype=np.float64)
        keep.append(yr)
        s.y_real, s.y_int = yr.ctypes.data, None
    else:
        yi = np.ascontiguousarray(y, dtype=np.int32)
        keep.append(yi)
        s.y_int, s.y_real = yi.ctypes.data, None
    if fam == BINOMIAL_LOGIT:
        ti = np.ascontiguousarray(trials, dtype=np.int32)
        keep.append(ti)
        s.trials = ti.ctypes.data
    s.n_classes = int(n_classes)
    s.G = int(G)
    if G:
        gi = np.ascontiguousarray(group, dtype=np.int32)
        keep.append(gi)
        s.group = gi.ctypes.data
    pri = dict(DEFAULT_PRIORS)
    pri.update(priors)
    for k, v in pri.items():
        setattr(s, k, float(v))
    return s, keep


def num_params(family, K, G=0, n_classes=0):
    fam = FAMILY[family] if isinstance(family, str) else int(family)
    if fam == ORDERED_LOGISTIC:
        return K + max(n_classes - 1, 0)
    if fam == CATEGORICAL_LOGIT:
        return n_classes * (1 + K)
    return (2 + G if G else 1) + K + (1 if fam in HAS_SCALE else 0)


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def build_port():
    subprocess.run(["make", "-C", HERE, "port"], check=True, capture_output=True)


class PortOracle:
    """Plain-C port (kind="port")."""
    kind = "port"

    def __init__(self, family, X, y, group=None, G=0, **priors):
        path = os.path.join(HERE, "_build", "libglm_oracle.so")
        if not os.path.exists(path):
            build_port()
        self.lib = C.CDLL(path)
        self.spec, self._keep = make_spec(family, X, y, group, G, **priors)
        self.P = self.lib.glm_oracle_num_params(C.byref(self.spec))

    def log_prob_grad(self, theta, propto=True, jacobian=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp = C.c_double()
        g = np.empty(self.P)
        err = C.create_string_buffer(512)
        rc = self.lib.glm_oracle_log_prob_grad(C.byref(self.spec), _dp(th), int(propto), int(jacobian),
                                               C.byref(lp), _dp(g), err, 512)
        if rc:
            raise OracleError(rc, err.value.decode())
        return lp.value, g

    def log_prob(self, theta, propto=False, jacobian=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp = C.c_double()
        err = C.create_string_buffer(512)
        rc = self.lib.glm_oracle_log_prob(C.byref(self.spec), _dp(th), int(propto), int(jacobian),
                                          C.byref(lp), err, 512)
        if rc:
            raise OracleError(rc, err.value.decode())
        return lp.value

    def leapfrog(self, eps, inv_metric, q, p, g, V):
        q, p, g = (np.array(a, dtype=np.float64) for a in (q, p, g))
        im = np.ascontiguousarray(inv_metric, dtype=np.float64)
        Vc = C.c_double(V)
        err = C.create_string_buffer(512)
        rc = self.lib.glm_oracle_leapfrog(C.byref(self.spec), C.c_double(eps), _dp(im), _dp(q), _dp(p), _dp(g),
                                          C.byref(Vc), err, 512)
        if rc:
            raise OracleError(rc, err.value.decode())
        return q, p, g, Vc.value


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_library_path():
    flags = _cpu_flags()
    v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
    v3 = {"avx2", "fma", "bmi2"} <= flags
    cands = (["v4"] if v4 else []) + (["v3"] if v3 else [])
    for isa in cands:
        p = os.path.join(HERE, "_ref", f"libref_oracle_{isa}.so")
        if os.path.exists(p):
            return p, isa
    return None, None


class RefOracle:
    """The compiled reference (kind="reference")."""
    kind = "reference"
    _lib = None
    _isa = None

    @classmethod
    def available(cls):
        return ref_library_path()[0] is not None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path, isa = ref_library_path()
            if path is None:
                raise RuntimeError("oracle/_ref/libref_oracle_*.so not built (make -C oracle ref needs /root/reference)")
            lib = C.CDLL(path)
            lib.ref_glm_create.restype = C.c_void_p
            lib.ref_glm_create.argtypes = [C.POINTER(GlmSpec)]
            lib.ref_glm_destroy.argtypes = [C.c_void_p]
            lib.ref_glm_num_params.argtypes = [C.c_void_p]
            for fn in (lib.ref_ess, lib.ref_rhat, lib.ref_mcse_mean, lib.ref_mcse_sd):
                fn.restype = C.c_double
                fn.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
            lib.ref_oracle_version.restype = C.c_char_p
            cls._lib, cls._isa = lib, isa
        return cls._lib

    @classmethod
    def glm_function(cls, family, X, y, alpha, beta, sigma=1.0, group=None, G=0, propto=True, operands_are_var=True,
                     sigma_is_var=True, trials=None):
        """The bare reference density stan::math::<family>_glm_lp*f<propto>(y, X, alpha | a[group], beta [, sigma])
        and its adjoints (ref_glm_function in oracle/ref/ref_oracle.cpp): returns lp, d_alpha, d_beta, d_sigma."""
        L = cls.lib()
        fam = FAMILY[family] if isinstance(family, str) else int(family)
        X = np.asfortranarray(X, dtype=np.float64)
        N, K = X.shape
        y = np.ascontiguousarray(y, dtype=np.float64 if fam == 2 else np.int32)
        a = np.ascontiguousarray(np.atleast_1d(alpha), dtype=np.float64)
        b = np.ascontiguousarray(beta, dtype=np.float64)
        grp = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        nt = None if trials is None else np.ascontiguousarray(trials, dtype=np.int32)
        lp, ds, err = C.c_double(), C.c_double(), C.create_string_buffer(1024)
        da, db = np.zeros_like(a), np.zeros(max(K, 1))
        ip = C.POINTER(C.c_int)
        rc = L.ref_glm_function(
            C.c_int(fam), C.c_int(int(propto)), C.c_int(int(operands_are_var)), C.c_int(int(sigma_is_var)),
            C.c_longlong(N), C.c_int(K), _dp(X), y.ctypes.data_as(ip) if fam != 2 else None,
            _dp(y) if fam == 2 else None, grp.ctypes.data_as(ip) if grp is not None else None, C.c_int(int(G)),
            _dp(a), _dp(b), C.c_double(float(sigma)), C.byref(lp), _dp(da), _dp(db), C.byref(ds), err, 1024,
            nt.ctypes.data_as(ip) if nt is not None else None)
        if rc:
            raise OracleError(rc, err.value.decode())
        return lp.value, da, db[:K], ds.value

    def __init__(self, family, X, y, group=None, G=0, **priors):
        self.L = self.lib()
        self.isa = self._isa
        spec, keep = make_spec(family, X, y, group, G, **priors)
        self.h = C.c_void_p(self.L.ref_glm_create(C.byref(spec)))
        del keep  # the reference model copies its data
        if not self.h:
            raise RuntimeError("ref_glm_create failed")
        self.P = self.L.ref_glm_num_params(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_glm_destroy(self.h)
            self.h = None

    def log_prob_grad(self, theta, propto=True, jacobian=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp = C.c_double()
        g = np.empty(self.P)
        err = C.create_string_buffer(1024)
        rc = self.L.ref_glm_log_prob_grad(self.h, _dp(th), int(propto), int(jacobian), C.byref(lp), _dp(g), err, 1024)
        if rc:
            raise OracleError(rc, err.value.decode())
        return lp.value, g

    def log_prob(self, theta, propto=False, jacobian=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp = C.c_double()
        err = C.create_string_buffer(1024)
        rc = self.L.ref_glm_log_prob(self.h, _dp(th), int(propto), int(jacobian), C.byref(lp), err, 1024)
        if rc:
            raise OracleError(rc, err.value.decode())
        return lp.value

    def gradient(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp = C.c_double()
        g = np.empty(self.P)
        err = C.create_string_buffer(1024)
        rc = self.L.ref_glm_gradient(self.h, _dp(th), C.byref(lp), _dp(g), err, 1024)
        if rc:
            raise OracleError(rc, err.value.decode())
        return lp.value, g

    def leapfrog(self, eps, inv_metric, q, p, g=None, V=0.0, init=False):
        q, p = (np.array(a, dtype=np.float64) for a in (q, p))
        g = np.zeros(self.P) if g is None else np.array(g, dtype=np.float64)
        im = np.ascontiguousarray(inv_metric, dtype=np.float64)
        Vc = C.c_double(V)
        err = C.create_string_buffer(1024)
        rc = self.L.ref_glm_leapfrog(self.h, C.c_double(eps), _dp(im), int(init), _dp(q), _dp(p), _dp(g),
                                     C.byref(Vc), err, 1024)
        if rc:
            raise OracleError(rc, err.value.decode())
        return q, p, g, Vc.value

    def _set_init_inv_metric(self, m):
        self.L.ref_set_init_inv_metric.argtypes = [C.POINTER(C.c_double), C.c_int]
        self.L.ref_set_init_inv_metric.restype = None
        if m is None:
            self.L.ref_set_init_inv_metric(None, 0)
        else:
            m = np.ascontiguousarray(m, dtype=np.float64)
            self.L.ref_set_init_inv_metric(_dp(m), int(m.size))

    def _set_adapt(self, adapt):
        """adapt: None (the services' defaults) or dict(gamma, kappa, t0, init_buffer, term_buffer, window)"""
        a = dict(gamma=0.05, kappa=0.75, t0=10.0, init_buffer=75, term_buffer=50, window=25)
        a.update(adapt or {})
        self.L.ref_set_adapt_params.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint, C.c_uint, C.c_uint]
        self.L.ref_set_adapt_params.restype = None
        self.L.ref_set_adapt_params(a["gamma"], a["kappa"], a["t0"], a["init_buffer"], a["term_buffer"], a["window"])

    def _set_jitter(self, j):
        self.L.ref_set_stepsize_jitter.argtypes = [C.c_double]
        self.L.ref_set_stepsize_jitter.restype = None
        self.L.ref_set_stepsize_jitter(float(j))

    def nuts(self, num_chains=4, seed=1, init_chain_id=1, init_radius=2.0, num_warmup=1000, num_samples=1000,
             stepsize=1.0, max_depth=10, delta=0.8, num_threads=0, stepsize_jitter=0.0, init_inv_metric=None, adapt=None):
        W = 7 + self.P
        self._set_jitter(stepsize_jitter)
        self._set_init_inv_metric(init_inv_metric)
        self._set_adapt(adapt)
        draws = np.empty((num_chains, num_warmup + num_samples, W))
        step = np.empty(num_chains)
        inv_metric = np.empty((num_chains, self.P))
        warm_lf = np.empty(num_chains)
        wall = C.c_double()
        err = C.create_string_buffer(2048)
        rc = self.L.ref_glm_nuts(self.h, num_chains, C.c_uint(seed), C.c_uint(init_chain_id), C.c_double(init_radius),
                                 num_warmup, num_samples, C.c_double(stepsize), max_depth, C.c_double(delta),
                                 num_threads, _dp(draws), _dp(step), _dp(inv_metric), _dp(warm_lf), C.byref(wall),
                                 err, 2048)
        if rc:
            raise OracleError(rc, err.value.decode())
        return dict(draws=draws[:, num_warmup:, :], warmup_draws=draws[:, :num_warmup, :], stepsize=step,
                    inv_metric=inv_metric, warm_leapfrogs=warm_lf, wall=wall.value)

    def nuts_device_host(self, num_chains=4, seed=1, init_chain_id=1, init_radius=2.0, num_warmup=1000, num_samples=1000,
                         stepsize=1.0, max_depth=10, delta=0.8, stepsize_jitter=0.0, init_inv_metric=None, adapt=None):
        """The PRODUCT's device-NUTS driver and per-chain state machine (stan_b200/cpp/b200/device_nuts.hpp,
        stan_b200/csrc/nuts_tree.cuh) built for the host over the reference's model and integrator
        (oracle/ref/nuts_host_backend.hpp): the CPU check of SURVEY 8f row 2.  Same outputs as nuts()."""
        W = 7 + self.P
        self._set_jitter(stepsize_jitter)
        self._set_init_inv_metric(init_inv_metric)
        self._set_adapt(adapt)
        draws = np.empty((num_chains, num_warmup + num_samples, W))
        step = np.empty(num_chains)
        inv_metric = np.empty((num_chains, self.P))
        stats = (C.c_long * 4)()
        err = C.create_string_buffer(2048)
        rc = self.L.ref_glm_nuts_device_host(self.h, num_chains, C.c_uint(seed), C.c_uint(init_chain_id),
                                             C.c_double(init_radius), num_warmup, num_samples, C.c_double(stepsize),
                                             max_depth, C.c_double(delta), _dp(draws), _dp(step), _dp(inv_metric), stats,
                                             err, 2048)
        if rc:
            raise OracleError(rc, err.value.decode())
        return dict(draws=draws[:, num_warmup:, :], warmup_draws=draws[:, :num_warmup, :], stepsize=step,
                    inv_metric=inv_metric, rounds=stats[0], lanes=stats[1], uniforms=stats[2], normal_vectors=stats[3])

    def nuts_transcript(self, which, num_chains=2, seed=1, init_chain_id=1, num_warmup=30, num_samples=20, num_thin=1,
                        save_warmup=True, refresh=0, stepsize=1.0, max_depth=10):
        """Everything the sample writer (S<chain>|), the diagnostic writer (D<chain>|) and the logger (L|) receive, as
        lines of text: which = 0 the reference's hmc_nuts_diag_e_adapt (chains one after the other), which = 1 the
        product's device-NUTS driver on the host backend."""
        buf = C.create_string_buffer(64 << 20)
        err = C.create_string_buffer(2048)
        rc = self.L.ref_glm_nuts_transcript(self.h, int(which), num_chains, C.c_uint(seed), C.c_uint(init_chain_id),
                                            num_warmup, num_samples, num_thin, int(save_warmup), refresh,
                                            C.c_double(stepsize), max_depth, buf, C.c_long(len(buf)), err, 2048)
        if rc:
            raise OracleError(rc, err.value.decode())
        return buf.value.decode().splitlines()

    # ---- stan::analyze ----
    @classmethod
    def _chains(cls, draws):
        d = np.asfortranarray(draws, dtype=np.float64)  # (n_draws, n_chains)
        return d, d.shape[0], d.shape[1]

    @classmethod
    def ess(cls, draws):
        d, n, c = cls._chains(draws)
        return cls.lib().ref_ess(_dp(d), n, c)

    @classmethod
    def rhat(cls, draws):
        d, n, c = cls._chains(draws)
        return cls.lib().ref_rhat(_dp(d), n, c)

    @classmethod
    def mcse_mean(cls, draws):
        d, n, c = cls._chains(draws)
        return cls.lib().ref_mcse_mean(_dp(d), n, c)

    @classmethod
    def mcse_sd(cls, draws):
        d, n, c = cls._chains(draws)
        return cls.lib().ref_mcse_sd(_dp(d), n, c)

    @classmethod
    def split_rank_normalized_ess(cls, draws):
        d, n, c = cls._chains(draws)
        b, t = C.c_double(), C.c_double()
        cls.lib().ref_split_rank_normalized_ess(_dp(d), n, c, C.byref(b), C.byref(t))
        return b.value, t.value
