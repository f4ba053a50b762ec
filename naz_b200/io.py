"""Wire / on-disk formats either side of the hot path, readable WITHOUT jax, numpyro, pyro or h5py.

What the reference writes and reads around the path (file:line):
  * posterior files  pickle({"params" [S, P] | pytree, "standard_params" [S, P], "scale" [S]})   bflow_jax_maf.py:333-335,
    :361-397 (arrays are jax Arrays) — `load_posterior` -> numpy / torch, straight into `FlowEngine.pack_draw_map` /
    `make_normalizing_flow(...)["lp_standard"]`;
  * checkpoint pair  (`checkpoint_file` = pickled numpyro HMCState, `posterior_file` as above)  :361-397 — the HMCState is
    opaque sampler state that only numpyro can resume from: `load_posterior` reads the posterior half, the other stays blocked
    on numpyro (stated, not emulated);
  * MLE flow  pickle(NormalizingFlow) with pyro modules inside   train_mle_all_data.py:77-78, calibrate.py:100-102 —
    `load_pickled_flow` recovers (params, param_shapes, masks, mask_skips, permutations), the tuple `torch_to_jax` returns,
    from the pickle's module tree without importing pyro or naz;
  * posterior-predictive draws  h5 dataset "ppds" [S, N, D]   calibrate.py:154-155 — `save_ppds` / `load_ppds`: .npy (always
    available, memory-mappable) or .h5 when h5py is importable; reading an existing .h5 stays blocked on h5py.
Nothing here touches the GPU: these are the callers' formats, the kernels see device tensors only.
"""
from __future__ import annotations

import io
import pickle
from collections import OrderedDict
from typing import Any, Dict, List, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# permissive unpickling: classes that cannot be imported become inert state holders
# ------------------------------------------------------------------------------------------------
class _Stub:
    """Placeholder for an instance of a class that is not importable here (pyro / naz / numpyro ...): keeps whatever
    state the pickle carries (`__dict__`, or the raw state under `_state`) and the original qualified name."""
    _qualname = "?"

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._state = state

    def __call__(self, *a, **k):      # a stubbed *function* used as a reduce callable
        return _Stub(*a, **k)


def _rebuild_jax_array(fun, args, arr_state, aval_state=None):
    """jax.Array.__reduce__ -> (_reconstruct_array, (fun, args, arr_state, aval_state)) where (fun, args, arr_state) is the
    __reduce__ triple of the equivalent numpy array (jax/_src/array.py): rebuild that numpy array and stop there."""
    a = fun(*args)
    a.__setstate__(arr_state)
    return a


class _PermissiveUnpickler(pickle.Unpickler):
    _SAFE_PREFIXES = ("numpy", "torch", "collections", "builtins", "copyreg", "functools", "_codecs")

    def find_class(self, module, name):
        if module.startswith("jax") and name == "_reconstruct_array":
            return _rebuild_jax_array
        if module.split(".")[0] in self._SAFE_PREFIXES:
            return super().find_class(module, name)
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (_Stub,), {"_qualname": f"{module}.{name}"})


def permissive_load(path_or_bytes) -> Any:
    if isinstance(path_or_bytes, (bytes, bytearray)):
        return _PermissiveUnpickler(io.BytesIO(path_or_bytes)).load()
    with open(path_or_bytes, "rb") as f:
        return _PermissiveUnpickler(f).load()


def _to_numpy(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    if isinstance(v, np.ndarray):
        return v
    if isinstance(v, (list, tuple)):
        return type(v)(_to_numpy(e) for e in v)
    if isinstance(v, dict):
        return {k: _to_numpy(e) for k, e in v.items()}
    return v


# ------------------------------------------------------------------------------------------------
# posterior files
# ------------------------------------------------------------------------------------------------
def load_posterior(path, device=None) -> Dict[str, Any]:
    """Posterior written by train_bayesian_flow* (pickle of jax arrays), by `save_posterior` (.npz), or a plain-numpy pickle.
    Returns {"standard_params": [S, P], "scale": [S] or [S, P], "params": [S, P] flat or the un-ravelled pytree} with whatever
    keys the file holds, as torch tensors (on `device` if given).  plot_svi.py:126 nests it under "posterior": unwrapped."""
    if str(path).endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        obj = {k: z[k] for k in z.files}
    else:
        obj = permissive_load(path)
    if isinstance(obj, dict) and "posterior" in obj and isinstance(obj["posterior"], dict):
        obj = obj["posterior"]
    if not isinstance(obj, dict):
        raise ValueError("not a posterior file: expected a dict of arrays")
    obj = _to_numpy(obj)

    def conv(v):
        if isinstance(v, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(v))
            t = t.float() if t.is_floating_point() else t
            return t.to(device) if device is not None else t
        if isinstance(v, (list, tuple)):
            return type(v)(conv(e) for e in v)
        return v
    return {k: conv(v) for k, v in obj.items()}


def save_posterior(path, posterior: Dict[str, Any]) -> None:
    """Portable .npz of the flat arrays ("standard_params", "scale", flat "params", ...); pytree-valued entries are
    ravelled in pytree order (the order `ravel_pytree` / `draw_params` use) under the same key."""
    out = {}
    for k, v in posterior.items():
        if isinstance(v, (list, tuple)):
            leaves = [np.asarray(_to_numpy(t)) for layer in v for pair in layer for t in pair]
            S = leaves[0].shape[0]
            out[k] = np.concatenate([l.reshape(S, -1) for l in leaves], axis=1)
        else:
            out[k] = np.asarray(_to_numpy(v))
    np.savez_compressed(path, **out)


# ------------------------------------------------------------------------------------------------
# pickled MLE flow (pyro modules inside) -> the torch_to_jax tuple
# ------------------------------------------------------------------------------------------------
def _children(mod) -> "OrderedDict[str, Any]":
    return getattr(mod, "_modules", None) or OrderedDict()


def _find_conditioners(root) -> List[Any]:
    """Depth-first walk of a (possibly stubbed) nn.Module tree: every module that has `layers` (ModuleList of masked
    linears) and a `permutation` buffer is one flow layer's autoregressive net, in registration (= flow) order."""
    found, seen = [], set()

    def visit(m):
        if id(m) in seen or m is None:
            return
        seen.add(id(m))
        mods = _children(m)
        bufs = getattr(m, "_buffers", None) or {}
        if "layers" in mods and ("permutation" in bufs or hasattr(m, "permutation")):
            found.append(m)
            return
        for c in mods.values():
            visit(c)
        for v in getattr(m, "__dict__", {}).values():       # plain attributes: flow_dist.transforms is a python list upstream
            if isinstance(v, (list, tuple)):
                for e in v:
                    if hasattr(e, "__dict__"):
                        visit(e)
            elif hasattr(v, "_modules") or isinstance(v, _Stub):
                visit(v)
    visit(root)
    return found


def flow_state_from_object(model) -> Tuple[list, list, list, list, list]:
    params, param_shapes, masks, mask_skips, perms = [], [], [], [], []
    for arn in _find_conditioners(model):
        bufs = getattr(arn, "_buffers", None) or {}
        layers = list(_children(_children(arn)["layers"]).values())
        lp, ls, lm = [], [], []
        for lin in layers:
            p = getattr(lin, "_parameters", {})
            W, b = p["weight"].detach().float(), p["bias"].detach().float()
            lp.append((W, b)); ls.append((tuple(W.shape), tuple(b.shape)))
            lm.append((getattr(lin, "_buffers", {}) or {}).get("mask"))
        params.append(lp); param_shapes.append(ls); masks.append([m.detach().float() for m in lm])
        ms = bufs.get("mask_skip", getattr(arn, "mask_skip", None))
        mask_skips.append(None if ms is None else ms.detach().float())
        pm = bufs.get("permutation", getattr(arn, "permutation", None))
        perms.append(pm.detach().long())
    if not params:
        raise ValueError("no autoregressive conditioners found in the pickled object")
    return params, param_shapes, masks, mask_skips, perms


def load_pickled_flow(path) -> Tuple[list, list, list, list, list]:
    """calibrate.py:100-102 (`model = pickle.load(pf); torch_to_jax(model)`) without pyro / naz importable: returns
    (params, param_shapes, masks, mask_skips, permutations) as torch CPU tensors."""
    return flow_state_from_object(permissive_load(path))


# ------------------------------------------------------------------------------------------------
# posterior-predictive draws
# ------------------------------------------------------------------------------------------------
def save_ppds(path, ppds, dataset: str = "ppds") -> str:
    """calibrate.py:154-155 writes h5 dataset "ppds" [S, N, D].  `.h5` needs h5py (RuntimeError naming the blocker when it is
    missing); any other suffix writes a .npy that `load_ppds` can memory-map."""
    arr = np.asarray(_to_numpy(ppds))
    path = str(path)
    if path.endswith(".h5") or path.endswith(".hdf5"):
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("writing .h5 needs h5py, which is not installed: use a .npy path") from e
        with h5py.File(path, "w") as hf:
            hf.create_dataset(dataset, data=arr)
        return path
    if not path.endswith(".npy"):
        path += ".npy"
    np.save(path, arr)
    return path


def load_ppds(path, dataset: str = "ppds", mmap: bool = True) -> np.ndarray:
    path = str(path)
    if path.endswith(".h5") or path.endswith(".hdf5"):
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("reading .h5 needs h5py, which is not installed") from e
        with h5py.File(path, "r") as hf:
            return hf[dataset][()]
    return np.load(path, mmap_mode="r" if mmap else None)
