"""Shared synthetic problems for the parity tests: the same seeded inputs are handed to the CPU oracle
(`oracle/`, test infrastructure) and to the device path (`sip_b200`)."""
import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix_uniform(seed: int, n: int) -> np.ndarray:
    """Counter-based splitmix64(seed + index) -> uniform(-1,1) float64 (SURVEY.md §8d generator)."""
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def synthetic_model(n, TF, seed=1234) -> np.ndarray:
    """v = 1500 + 3000*depth + 300*sin(6 pi i/n1)*cos(4 pi j/n2) + 150*u  (depth = slowest axis)."""
    n = tuple(int(v) for v in n)
    N = int(np.prod(n))
    u = splitmix_uniform(seed, N).reshape(n, order="F")
    if len(n) == 2:
        i = np.arange(n[0])[:, None]
        k = np.arange(n[1])[None, :]
        v = 1500.0 + 3000.0 * k / (n[1] - 1) + 300.0 * np.sin(6 * np.pi * i / n[0]) + 150.0 * u
    else:
        i = np.arange(n[0])[:, None, None]
        j = np.arange(n[1])[None, :, None]
        k = np.arange(n[2])[None, None, :]
        v = 1500.0 + 3000.0 * k / (n[2] - 1) + 300.0 * np.sin(6 * np.pi * i / n[0]) * np.cos(4 * np.pi * j / n[1]) + 150.0 * u
    return np.ascontiguousarray(v.ravel(order="F").astype(TF))


def tv_l1(n, d, TF, m) -> float:
    """||TV m||_1 computed in float64 with plain differences (only used to size the l1 ball)."""
    x = m.astype(np.float64).reshape(n, order="F")
    return float(sum(np.abs(np.diff(x, axis=a)).sum() / float(TF(d[a])) for a in range(len(n))))


def spec_config1(n=(64, 64), TF=np.float64):
    """bounds ∩ TV-l1 ∩ D_z slope bounds (BASELINE config 1: projection_intersection_2D.jl-style)."""
    d = (25.0, 6.0)
    m = synthetic_model(n, TF)
    sets = [("bounds", "identity", 1500.0, 4500.0), ("l1", "TV", 0.0, 0.5 * tv_l1(n, d, TF, m)),
            ("bounds", "D_z", 0.0, 1e6)]
    return dict(n=n, d=d, TF=TF, m=m, sets=sets, mode="matrix")


def spec_config2(n=(24, 24, 24), TF=np.float32):
    """bounds ∩ anisotropic TV ∩ lateral smoothness (BASELINE config 2: test_scaling_3D-style)."""
    d = (25.0, 25.0, 25.0)
    m = synthetic_model(n, TF)
    sets = [("bounds", "identity", 1500.0, 6000.0), ("l1", "TV", 0.0, 0.5 * tv_l1(n, d, TF, m)),
            ("bounds", "D_x", -1.0, 1.0), ("bounds", "D_y", -1.0, 1.0)]
    return dict(n=n, d=d, TF=TF, m=m, sets=sets, mode="tensor")


def spec_config3(n=(24, 24, 24), TF=np.float32, frac=0.05):
    """bounds ∩ TV-l1 ∩ cardinality of the discrete gradient (BASELINE config 3)."""
    d = (25.0, 25.0, 25.0)
    m = synthetic_model(n, TF)
    M = sum(int(np.prod([v - 1 if a == b else v for b, v in enumerate(n)])) for a in range(3))
    sets = [("bounds", "identity", 1500.0, 6000.0), ("l1", "TV", 0.0, 0.5 * tv_l1(n, d, TF, m)),
            ("cardinality", "TV", 0, int(frac * M))]
    return dict(n=n, d=d, TF=TF, m=m, sets=sets, mode="tensor")


def spec_config4(n=(24, 24, 24), TF=np.float32):
    """constraints of examples/test_scaling_3D.jl:41-74 (multilevel config 4)."""
    d = (25.0, 25.0, 25.0)
    m = synthetic_model(n, TF)
    sets = [("bounds", "identity", 1500.0, 6000.0), ("bounds", "D_z", 0.0, 1e6), ("bounds", "D_x", -1.0, 1.0),
            ("bounds", "D_y", -1.0, 1.0)]
    return dict(n=n, d=d, TF=TF, m=m, sets=sets, mode="tensor")


def build(api, spec, options=None, types=None):
    """Run the reference call sequence setup_constraints -> PARSDMM_precompute_distribute with `api`
    (the oracle modules or the sip_b200 package).  `types` supplies set_definitions/compgrid/options."""
    T = types or api
    cg = T.compgrid(tuple(spec["d"]), tuple(spec["n"]))
    cons = [T.set_definitions(st, op, lo, hi, (spec["mode"], "")) for (st, op, lo, hi) in spec["sets"]]
    opt = options if options is not None else T.PARSDMM_options()
    opt.FL = spec["TF"]
    P_sub, TD_OP, set_Prop = api.setup_constraints(cons, cg, spec["TF"])
    TD_OP, AtA, l, y = api.PARSDMM_precompute_distribute(TD_OP, set_Prop, cg, opt)
    return dict(cg=cg, cons=cons, opt=opt, P_sub=P_sub, TD_OP=TD_OP, set_Prop=set_Prop, AtA=AtA, l=l, y=y)


class OracleAPI:
    """Adapter giving the oracle modules the same flat namespace as sip_b200."""

    def __init__(self):
        from oracle import operators, parsdmm, projectors, setup, sip_types
        self.compgrid = sip_types.compgrid
        self.set_definitions = sip_types.set_definitions
        self.PARSDMM_options = sip_types.PARSDMM_options
        self.setup_constraints = setup.setup_constraints
        self.PARSDMM_precompute_distribute = setup.PARSDMM_precompute_distribute
        self.PARSDMM_precompute_distribute_Minkowski = setup.PARSDMM_precompute_distribute_Minkowski
        self.PARSDMM = parsdmm.PARSDMM
        self.get_TD_operator = operators.get_TD_operator
        self.mat2CDS = operators.mat2CDS
        self.ops = operators
        self.proj = projectors
        self.parsdmm = parsdmm
        self.types = sip_types


def build_minkowski(api, n=(32, 28), TF=np.float32, options=None):
    """BASELINE config 5 (examples/GeneralizedMinkowski/example_2D_Minkowski_projection.jl:61-137):
    m ≈ x1 + x2 with  c1: vector bounds (water layer) + D_z >= 0;  c2: bounds + TV-l1(0.15);  sum: bounds."""
    d = (25.0, 6.0)
    m = synthetic_model(n, TF)
    cg = api.compgrid(d, n)
    lo = np.full(n, 1500.0, dtype=TF)
    hi = np.full(n, 4500.0, dtype=TF)
    hi[:, : max(n[1] // 8, 1)] = 1500.0
    mode = ("matrix", "")
    c1 = [api.set_definitions("bounds", "identity", lo.ravel(order="F"), hi.ravel(order="F"), mode),
          api.set_definitions("bounds", "D_z", 0.0, 1e6, mode)]
    c2 = [api.set_definitions("bounds", "identity", -1500.0, 1500.0, mode),
          api.set_definitions("l1", "TV", 0.0, 0.15 * tv_l1(n, d, TF, m), mode)]
    cs = [api.set_definitions("bounds", "identity", 1500.0, 4500.0, mode)]
    P1, T1, S1 = api.setup_constraints(c1, cg, TF)
    P2, T2, S2 = api.setup_constraints(c2, cg, TF)
    P3, T3, S3 = api.setup_constraints(cs, cg, TF)
    opt = options if options is not None else api.PARSDMM_options()
    opt.FL = TF
    opt.Minkowski = True
    opt.feas_tol, opt.obj_tol, opt.evol_rel_tol = 1e-3, 1e-3, 1e-5
    TD_OP, set_Prop, AtA, l, y = api.PARSDMM_precompute_distribute_Minkowski(T1, T2, T3, S1, S2, S3, cg, opt)
    return dict(cg=cg, opt=opt, P_sub=list(P1) + list(P2) + list(P3), TD_OP=TD_OP, set_Prop=set_Prop, AtA=AtA, l=l, y=y, m=m)
