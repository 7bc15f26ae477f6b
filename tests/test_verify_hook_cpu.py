"""tools/verify_against_ref.py cannot meet the real reference in this image (no jax / ksim / xax / equinox /
distrax), so its plumbing is exercised against STUB modules: dump -> tests/golden-style ref_*.npz -> check, a deliberate
mismatch is caught and named, and without the packages the hook reports "unverified" instead of failing.  The stubs restate
the oracle's own formulas, so a pass here says nothing about parity -- only that the hook works when the day comes."""

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import pytest

import kbot_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def _load_hook():
    spec = importlib.util.spec_from_file_location("verify_against_ref", ROOT / "tools" / "verify_against_ref.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub_modules(lpf_form="rc"):
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    jnp.asarray, jnp.zeros = np.asarray, np.zeros
    jnp.clip = lambda x, min=None, max=None: np.clip(x, min, max)
    jax.numpy = jnp
    jax.vmap = lambda f: (lambda *xs: np.stack([np.asarray(f(*r)) for r in zip(*xs)]))
    jax.random = types.SimpleNamespace(PRNGKey=lambda s: np.array([0, s], np.uint32))
    jax.nn = types.SimpleNamespace(softplus=O.softplus)
    jax.tree_util = types.SimpleNamespace(tree_leaves=lambda t: [t.y] if hasattr(t, "y") else list(t or []))
    jax.__version__ = "stub"

    xax = types.ModuleType("xax")
    xax.quat_to_euler = lambda q: O.quat_to_euler(np.asarray(q))
    xax.euler_to_quat = lambda e: O.euler_to_quat(np.asarray(e))
    xax.rotate_vector_by_quat = lambda v, q, inverse=False: O.rotate_vector_by_quat(np.asarray(v), np.asarray(q), inverse=inverse)
    xax.get_norm = lambda x, kind: np.asarray(x) ** 2
    xax.__version__ = "stub"

    class LowPassFilterParams:
        def __init__(self, y):
            self.y = y

        @classmethod
        def initialize(cls, n):
            return cls(np.zeros(n, np.float32))

    def lowpass_one_pole(x, dt, fc, params):
        alpha = np.float32(O.OracleParams(lpf_form=lpf_form, ctrl_dt=float(dt), cutoff_frequency=float(fc)).lpf_alpha)
        y = params.y + alpha * (np.asarray(x) - params.y)
        return y, LowPassFilterParams(y)

    ksim = types.ModuleType("ksim")
    ksim.LowPassFilterParams, ksim.lowpass_one_pole = LowPassFilterParams, lowpass_one_pole
    ksim.__version__ = "stub"

    class LSTMCell:
        def __init__(self, i, h, key):
            r = np.random.default_rng(int(key[1]))
            self.weight_ih = r.normal(0, 0.2, (4 * h, i)).astype(np.float32)
            self.weight_hh = r.normal(0, 0.2, (4 * h, h)).astype(np.float32)
            self.bias = r.normal(0, 0.2, (4 * h,)).astype(np.float32)

        def __call__(self, x, hc):
            h, c = O.lstm_cell(self.weight_ih, self.weight_hh, self.bias, x[None], hc[0][None], hc[1][None])
            return h[0], c[0]

    class Linear:
        def __init__(self, i, o, key):
            r = np.random.default_rng(int(key[1]))
            self.weight, self.bias = r.normal(0, 0.3, (o, i)).astype(np.float32), r.normal(0, 0.3, (o,)).astype(np.float32)

        def __call__(self, x):
            return O.linear(self.weight, self.bias, x[None])[0]

    eqx = types.ModuleType("equinox")
    eqx.nn = types.SimpleNamespace(LSTMCell=LSTMCell, Linear=Linear)
    eqx.__version__ = "stub"

    class MultivariateNormalDiag:
        def __init__(self, loc, scale_diag):
            self.loc, self.scale = np.asarray(loc), np.asarray(scale_diag)

        def log_prob(self, a):
            return O.mvn_log_prob(self.loc, self.scale, np.asarray(a))

        def entropy(self):
            return O.mvn_entropy(self.scale)

        def mode(self):
            return self.loc

        def stddev(self):
            return self.scale

        def sample(self, seed):
            return self.loc + self.scale * np.random.default_rng(int(seed[1])).standard_normal(self.loc.shape).astype(np.float32)

    distrax = types.ModuleType("distrax")
    distrax.MultivariateNormalDiag = MultivariateNormalDiag
    distrax.__version__ = "stub"
    return {"jax": jax, "jax.numpy": jnp, "xax": xax, "ksim": ksim, "equinox": eqx, "distrax": distrax}


SELF_CONTAINED = ["quat_helpers", "lowpass_one_pole", "lstm_cell", "mvn_diag", "softplus"]


def test_hook_reports_unverified_without_the_reference_packages(tmp_path):
    V = _load_hook()
    ok, why = V.reference_available()
    if ok:
        pytest.skip("the reference's packages ARE importable here: run tools/verify_against_ref.py --dump and commit the goldens")
    assert isinstance(why, str) and why
    assert V.main(["--dump", "--dir", str(tmp_path)]) == 3                  # nothing dumped, no exception
    rep = V.check(tmp_path)
    assert set(rep) == set(V.ITEMS) and all(v["status"] == "unverified" for v in rep.values())
    assert set(V.DUMPERS) == set(V.ITEMS)                                   # every [U] item has a dumper


def test_hook_dump_and_check_round_trip_with_stubbed_packages(tmp_path, monkeypatch):
    V = _load_hook()
    for name, mod in _stub_modules().items():
        monkeypatch.setitem(sys.modules, name, mod)
    rep = V.dump(tmp_path, only=SELF_CONTAINED)
    assert rep["available"] and all(rep["items"][k] == "dumped" for k in SELF_CONTAINED), rep
    chk = V.check(tmp_path)
    for k in SELF_CONTAINED:
        assert chk[k]["status"] == "verified", (k, chk[k])
    assert chk["compute_ppo_inputs"]["status"] == "unverified"
    # items that need the real Task (physics model, ksim runtime) fail softly under the stubs: reported, not raised
    rep2 = V.dump(tmp_path, only=["not_upright"])
    assert rep2["items"]["not_upright"].startswith("FAILED")


def test_hook_names_the_matching_form_on_a_mismatch(tmp_path, monkeypatch):
    """If the fork's low-pass filter used the other coefficient form, --check says so and names it (a one-line oracle fix)."""
    V = _load_hook()
    for name, mod in _stub_modules(lpf_form="exp").items():
        monkeypatch.setitem(sys.modules, name, mod)
    V.dump(tmp_path, only=["lowpass_one_pole"])
    chk = V.check(tmp_path)
    assert chk["lowpass_one_pole"]["status"] == "MISMATCH" and "exp" in chk["lowpass_one_pole"]["detail"]
    assert V.main(["--check", "--dir", str(tmp_path)]) == 1
