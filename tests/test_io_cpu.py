"""Wire / on-disk formats (naz_b200/io.py) — read without jax, numpyro, pyro or h5py."""
import io
import pickle
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from naz_b200 import io as nio


def _fake_jax_array_pickle(obj):
    """Bytes of a pickle in which every numpy array inside `obj` is encoded the way jax.Array pickles itself:
    reduce -> jax._src.array._reconstruct_array(fun, args, arr_state, aval_state) (jax/_src/array.py)."""
    mod = types.ModuleType("jax._src.array")

    def _reconstruct_array(fun, args, arr_state, aval_state):   # never called here: loading goes through naz_b200.io
        raise AssertionError
    _reconstruct_array.__module__, _reconstruct_array.__qualname__ = "jax._src.array", "_reconstruct_array"
    mod._reconstruct_array = _reconstruct_array

    class FakeJaxArray:
        def __init__(self, a):
            self.a = np.asarray(a)

        def __reduce__(self):
            fun, args, state = self.a.__reduce__()
            return (_reconstruct_array, (fun, args, state, {"weak_type": False, "named_shape": {}}))

    def wrap(v):
        if isinstance(v, np.ndarray):
            return FakeJaxArray(v)
        if isinstance(v, dict):
            return {k: wrap(e) for k, e in v.items()}
        if isinstance(v, (list, tuple)):
            return type(v)(wrap(e) for e in v)
        return v
    saved = {k: sys.modules.get(k) for k in ("jax", "jax._src", "jax._src.array")}
    sys.modules.update({"jax": types.ModuleType("jax"), "jax._src": types.ModuleType("jax._src"), "jax._src.array": mod})
    try:
        return pickle.dumps(wrap(obj))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_posterior_file_with_jax_arrays_loads_without_jax(tmp_path):
    rng = np.random.default_rng(0)
    S, P = 7, 33
    post = {"standard_params": rng.uniform(-1, 1, (S, P)).astype(np.float32), "scale": np.full(S, 0.25, np.float32),
            "params": rng.normal(size=(S, P)).astype(np.float32)}
    f = tmp_path / "posterior.pkl"
    f.write_bytes(_fake_jax_array_pickle(post))
    assert "jax" not in sys.modules
    with pytest.raises(Exception):
        pickle.loads(f.read_bytes())                       # the stock unpickler needs jax
    got = nio.load_posterior(f)
    assert set(got) == set(post)
    for k in post:
        assert isinstance(got[k], torch.Tensor) and got[k].dtype == torch.float32 and np.array_equal(got[k].numpy(), post[k])
    # plot_svi.py:126: nested under "posterior"
    g = tmp_path / "svi.pkl"
    g.write_bytes(_fake_jax_array_pickle({"posterior": post, "svi_params": {"mu_param_q": post["params"][0]}}))
    assert np.array_equal(nio.load_posterior(g)["standard_params"].numpy(), post["standard_params"])
    # portable npz round trip, pytree-valued "params" ravelled in pytree order
    from naz_b200.flows.bflow_maf import ravel_pytree
    best = [[(torch.randn(5, 3), torch.randn(5)), (torch.randn(4, 5), torch.randn(4))]]
    flat, unravel = ravel_pytree(best)
    batch = torch.stack([flat, 2 * flat, 3 * flat])
    nio.save_posterior(tmp_path / "p.npz", {"params": unravel(batch), "scale": torch.ones(3)})
    back = nio.load_posterior(tmp_path / "p.npz")
    assert torch.equal(back["params"], batch) and back["scale"].shape == (3,)


def test_pickled_pyro_flow_is_read_without_pyro(tmp_path):
    """A pickle whose classes live in modules that do not exist here (pyro.nn..., naz.flows.flow) still yields the
    torch_to_jax tuple: weights, masks, mask_skip and permutation of every flow layer, in flow order."""
    from naz_b200.flows.made import create_mask
    names = ["pyro", "pyro.nn", "pyro.nn.auto_reg_nn", "pyro.distributions", "pyro.distributions.transforms", "naz", "naz.flows", "naz.flows.flow"]
    mods = {n: types.ModuleType(n) for n in names}

    class MaskedLinear(nn.Linear):
        def __init__(self, i, o, mask):
            super().__init__(i, o)
            self.register_buffer("mask", mask)
    MaskedLinear.__module__, MaskedLinear.__qualname__ = "pyro.nn.auto_reg_nn", "MaskedLinear"

    class ConditionalAutoRegressiveNN(nn.Module):
        def __init__(self, D, C, hidden, perm):
            super().__init__()
            self.masks, self.mask_skip = create_mask(D, C, hidden, perm, 2)
            self.register_buffer("permutation", perm)
            dims = [D + C] + hidden + [2 * D]
            self.layers = nn.ModuleList([MaskedLinear(dims[j], dims[j + 1], self.masks[j]) for j in range(len(dims) - 1)])
    ConditionalAutoRegressiveNN.__module__, ConditionalAutoRegressiveNN.__qualname__ = "pyro.nn.auto_reg_nn", "ConditionalAutoRegressiveNN"

    class ConditionalAffineAutoregressive(nn.Module):
        def __init__(self, arn):
            super().__init__()
            self.nn = arn
    ConditionalAffineAutoregressive.__module__ = "pyro.distributions.transforms"
    ConditionalAffineAutoregressive.__qualname__ = "ConditionalAffineAutoregressive"

    class NormalizingFlow(nn.Module):
        def __init__(self, D, C, hidden, L):
            super().__init__()
            self.embedding_net = nn.Identity()
            self.nets = [ConditionalAutoRegressiveNN(D, C, hidden, torch.randperm(D)) for _ in range(L)]
            self.transforms = [ConditionalAffineAutoregressive(a) for a in self.nets]
            self.flow = nn.ModuleList(self.transforms)
    NormalizingFlow.__module__, NormalizingFlow.__qualname__ = "naz.flows.flow", "NormalizingFlow"
    mods["pyro.nn.auto_reg_nn"].MaskedLinear = MaskedLinear
    mods["pyro.nn.auto_reg_nn"].ConditionalAutoRegressiveNN = ConditionalAutoRegressiveNN
    mods["pyro.distributions.transforms"].ConditionalAffineAutoregressive = ConditionalAffineAutoregressive
    mods["naz.flows.flow"].NormalizingFlow = NormalizingFlow
    torch.manual_seed(3)
    flow = NormalizingFlow(3, 2, [16, 12], 4)
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(mods)
    try:
        blob = pickle.dumps(flow)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    f = tmp_path / "mle_flow.pkl"
    f.write_bytes(blob)
    with pytest.raises(Exception):
        pickle.loads(blob)                                 # stock unpickler: No module named 'naz' / 'pyro'
    params, shapes, masks, mask_skips, perms = nio.load_pickled_flow(f)
    assert len(params) == 4 and [len(p) for p in params] == [3] * 4
    for l, arn in enumerate(flow.nets):
        assert torch.equal(perms[l], arn.permutation) and torch.equal(mask_skips[l], arn.mask_skip)
        for j, lin in enumerate(arn.layers):
            assert torch.equal(params[l][j][0], lin.weight.detach()) and torch.equal(params[l][j][1], lin.bias.detach())
            assert torch.equal(masks[l][j], lin.mask) and shapes[l][j] == (tuple(lin.weight.shape), tuple(lin.bias.shape))


def test_ppd_writer_round_trip_and_h5_blocker(tmp_path):
    ppds = np.random.default_rng(1).normal(size=(5, 40, 2)).astype(np.float32)
    p = nio.save_ppds(tmp_path / "ppds_label", torch.from_numpy(ppds))
    assert p.endswith(".npy")
    back = nio.load_ppds(p)
    assert back.shape == (5, 40, 2) and np.array_equal(np.asarray(back), ppds)
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="h5py"):
            nio.save_ppds(tmp_path / "ppds.h5", ppds)
        with pytest.raises(RuntimeError, match="h5py"):
            nio.load_ppds(tmp_path / "ppds.h5")


def test_flow_state_of_our_own_flow_equals_torch_to_jax(tmp_path):
    """The same walker reads a pickled naz_b200 flow (classes importable): identical to torch_to_jax on the live object."""
    from naz_b200.flows import NormalizingFlow
    from naz_b200.flows.bflow_maf import torch_to_jax
    torch.manual_seed(5)
    flow = NormalizingFlow("maf", None, 3, 2, [16, 12], 3)
    f = tmp_path / "flow.pkl"
    with open(f, "wb") as pf:
        pickle.dump(flow, pf)
    params, shapes, masks, mask_skips, perms = nio.load_pickled_flow(f)
    p2, s2, m2, ms2, pm2 = torch_to_jax(flow)
    assert shapes == s2
    for l in range(3):
        assert torch.equal(perms[l], pm2[l]) and torch.equal(mask_skips[l], ms2[l])
        for j in range(3):
            assert torch.equal(params[l][j][0], p2[l][j][0]) and torch.equal(params[l][j][1], p2[l][j][1]) and torch.equal(masks[l][j], m2[l][j])
