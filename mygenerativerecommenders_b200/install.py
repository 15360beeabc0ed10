"""Drop-in installation into a reference checkout.

Two independent switches, both reversible:

* :func:`install_ops` rebinds the nine functions of the reference's operator layer
  ``generative_recommenders_pl.models.utils.ops`` (ops.py:18,41,67,117,149,171,190,210,229) to
  this package's kernels.  Every caller in the reference reaches them through ``ops.<fn>``
  attribute access (hstu.py:179-204,502,512,661; generative_recommenders.py:384,407-424;
  retrieval.py:30; ranking.py:33; candidate_index.py:80-88), so nothing else changes — this is
  the slot the reference reserves for ``torch.ops.fbgemm.*`` (ops.py:27,51,87).
* :func:`install_modules` swaps the Hydra ``_target_`` classes (configs/model/hstu.yaml:18,41,44,
  49,52,55) for the fused B200 modules by rebinding the class attributes of the reference's
  modules.  The cleaner way is to point ``_target_`` at this package: see ``configs/`` here.
"""
from __future__ import annotations

import importlib
from typing import Dict, Tuple

from . import candidate_index, hstu, losses, negative_sampler, ops, similarity, top_k

_REF = "generative_recommenders_pl.models"
_saved: Dict[Tuple[str, str], object] = {}

_MODULE_MAP = {
    f"{_REF}.sequential_encoders.hstu": {
        "HSTU": hstu.HSTU, "HSTUJagged": hstu.HSTUJagged,
        "SequentialTransductionUnitJagged": hstu.SequentialTransductionUnitJagged,
        "RelativeBucketedTimeAndPositionBasedBias": hstu.RelativeBucketedTimeAndPositionBasedBias,
    },
    f"{_REF}.indexing.top_k": {"MIPSBruteForceTopK": top_k.MIPSBruteForceTopK},
    f"{_REF}.indexing.candidate_index": {"CandidateIndex": candidate_index.CandidateIndex},
    f"{_REF}.negatives_samples.negative_sampler": {
        "LocalNegativesSampler": negative_sampler.LocalNegativesSampler,
        "InBatchNegativesSampler": negative_sampler.InBatchNegativesSampler,
    },
    f"{_REF}.similarity.dot_product": {"DotProductSimilarity": similarity.DotProductSimilarity},
    f"{_REF}.losses.autoregressive_losses": {"SampledSoftmaxLoss": losses.SampledSoftmaxLoss},
}


def _rebind(mod_name: str, attr: str, value) -> None:
    mod = importlib.import_module(mod_name)
    _saved.setdefault((mod_name, attr), getattr(mod, attr))
    setattr(mod, attr, value)


def install_ops() -> None:
    for name in ops.__all__:
        _rebind(f"{_REF}.utils.ops", name, getattr(ops, name))


def install_modules() -> None:
    import sys
    for mod_name, mapping in _MODULE_MAP.items():
        for attr, cls in mapping.items():
            _rebind(mod_name, attr, cls)
    # modules that did ``from ... import InBatchNegativesSampler`` before this call hold their own
    # binding (models/retrieval.py:104 isinstance-checks against it): rebind those too, without
    # importing them (they pull in lightning / hydra)
    flat = {attr: cls for mapping in _MODULE_MAP.values() for attr, cls in mapping.items()}
    for name in (f"{_REF}.retrieval", f"{_REF}.ranking", f"{_REF}.generative_recommenders"):
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for attr, cls in flat.items():
            if hasattr(mod, attr):
                _rebind(name, attr, cls)


def install() -> None:
    install_ops()
    install_modules()


def uninstall() -> None:
    for (mod_name, attr), value in list(_saved.items()):
        setattr(importlib.import_module(mod_name), attr, value)
    _saved.clear()
