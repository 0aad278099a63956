"""DistributedHotPath (proteuscfd_b200/parallel.py) = SolutionSpace::NewtonIterate across ranks.  With a recording stand-in
for the GPU context and the exchange, the order of phase calls and halo exchanges is checked against the reference's
own sequence (ucs/solutionSpace.tcc:640-904; gradient.tcc:98; limiters.tcc:128; crs.tcc:88,146) -- host logic only, no
GPU, no numerics."""
from proteuscfd_b200 import capi
from proteuscfd_b200.parallel import DistributedHotPath

NAMES = {capi.F_Q: "q", capi.F_QGRAD: "qgrad", capi.F_LIMITER: "limiter", capi.F_X: "x", capi.F_LSQ_S: "s", capi.F_LSQ_SW: "sw",
         capi.F_TVAR: "tvar", capi.F_TGRAD: "tgrad", capi.F_TURB_X: "turb_x"}


class Recorder:
    """stands in for capi.Context and for an exchange: every call is appended to one shared log"""

    def __init__(self, log, neqn=5, clip_hits=()):
        self.log, self.neqn, self.clip_hits, self.nfused = log, neqn, list(clip_hits), 0

    def update(self, field):
        self.log.append("halo:" + NAMES[field])

    def residual_fused(self, want_norms=False):
        self.log.append("residual_fused")
        hit = self.clip_hits[self.nfused] if self.nfused < len(self.clip_hits) else False
        self.nfused += 1
        return None, hit

    def turb_phase(self, phase, want_norm=False):
        self.log.append(f"turb_phase{phase}")
        return 2.0 if want_norm else None

    def __getattr__(self, name):
        def call(*a, **k):
            self.log.append(name)
        return call


def run(fused, implicit, neqn=5, clip_hits=(), nsgs=2, refresh=True):
    log = []
    ctx = Recorder(log, neqn, clip_hits)
    hp = DistributedHotPath(ctx, ctx, any_rank=(lambda hit: hit) if fused else None)
    hp.setup()
    if implicit:
        hp.implicit_iterate(nsgs, refresh_jac=refresh)
    else:
        hp.explicit_iterate(refresh_dt=True)
    return log, hp


SETUP = ["lsq_coefficients", "halo:s", "halo:sw"]
HEAD = ["update_bcs", "halo:q", "gradient", "halo:qgrad"]


def test_explicit_iteration_follows_newton_iterate():
    log, _ = run(fused=False, implicit=False)
    assert log == SETUP + ["timestep"] + HEAD + ["limiter", "halo:limiter", "residual", "explicit_solve", "halo:q"]


def test_implicit_iteration_follows_newton_iterate():
    log, _ = run(fused=False, implicit=True, nsgs=3)
    assert log == (SETUP + ["timestep", "jacobian"] + HEAD + ["limiter", "halo:limiter", "residual", "prepare_sgs", "blank_x", "halo:x"]
                   + ["sgs", "halo:x"] * 3 + ["apply_dq", "halo:q"])
    log, _ = run(fused=False, implicit=True, nsgs=1, refresh=False)      # Jacobian kept: no time step, no assembly
    assert "jacobian" not in log and "timestep" not in log


def test_fused_pair_exchanges_the_raw_limiter_once():
    for neqn in (5, 9):        # both eqnset families take the fused path
        log, hp = run(fused=True, implicit=False, neqn=neqn)
        assert log == SETUP + ["timestep"] + HEAD + ["limiter_raw", "halo:limiter", "residual_fused", "explicit_solve", "halo:q"]
        assert hp.clip_fallbacks == 0


def test_clip_hit_on_any_rank_falls_back_to_the_ordered_path():
    log, hp = run(fused=True, implicit=True, clip_hits=[True], nsgs=1)
    i = log.index("residual_fused")
    assert log[i + 1: i + 4] == ["limiter", "halo:limiter", "residual"]
    assert hp.clip_fallbacks == 1 and log.count("gradient") == 1       # the gradient is not redone


def test_turbulence_compute_follows_the_reference_exchange_points():
    """TurbulenceModel::Compute (ucs/turb.tcc:163-339): UpdateBCs, halo of tvar (:185), gradient + its halo (:192,
    gradient.tcc:98), assembly and the all-reduced residual (:256), nSgs sweeps with a halo of x each (crs.tcc:146), the
    update and the halo of tvar (:325), then the eddy viscosity"""
    log = []
    ctx = Recorder(log)
    hp = DistributedHotPath(ctx, ctx, allreduce_sum=lambda v: 3.0 * v)
    s = hp.turb_compute(2, want_norm=True)
    assert log == ["turb_phase0", "halo:tvar", "turb_phase1", "halo:tgrad", "turb_phase2", "turb_phase3", "halo:turb_x",
                   "turb_phase3", "halo:turb_x", "turb_phase4", "halo:tvar", "turb_phase5"]
    assert s == 6.0      # the residual sum of squares goes through the all-reduce
    log.clear()
    assert hp.turb_compute(1) is None and log.count("turb_phase3") == 1
