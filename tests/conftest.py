import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return torch.load(GOLDEN / f"{name}.pt", map_location="cpu", weights_only=True)
    return load


def hstu_case(g, name):
    """Unpack one case of tests/golden/hstu.pt."""
    B, max_seq, out_len, D, H, dqk, dv, blocks = [int(v) for v in g[f"{name}.cfg"]]
    sd = {k[len(name) + 4:]: v for k, v in g.items() if k.startswith(f"{name}.sd.")}
    grads = {k[len(name) + 6:]: v for k, v in g.items() if k.startswith(f"{name}.grad.")}
    return dict(B=B, max_seq=max_seq, out_len=out_len, N=max_seq + out_len, D=D, H=H, dqk=dqk,
                dv=dv, blocks=blocks, sd=sd, grads=grads, lengths=g[f"{name}.lengths"],
                ts=g[f"{name}.ts"], x=g[f"{name}.x"], w=g[f"{name}.w"], y=g[f"{name}.y"],
                dx=g[f"{name}.dx"])


def hstu_incremental_case(g, name):
    """Unpack one case of tests/golden/hstu_incremental.pt (oracle/make_golden.py)."""
    B, max_seq, out_len, D, H, dqk, dv, blocks = [int(v) for v in g[f"{name}.cfg"]]
    sd = {k[len(name) + 4:]: v for k, v in g.items() if k.startswith(f"{name}.sd.")}
    names = ("v", "padded_q", "padded_k", "out")
    cache0 = [tuple(g[f"{name}.cache0.{i}.{nm}"] for nm in names) for i in range(blocks)]
    cache1 = None
    if f"{name}.cache1.0.v" in g:
        cache1 = [tuple(g[f"{name}.cache1.{i}.{nm}"] for nm in names) for i in range(blocks)]
    return dict(B=B, max_seq=max_seq, out_len=out_len, N=max_seq + out_len, D=D, H=H, dqk=dqk,
                dv=dv, blocks=blocks, sd=sd, lengths=g[f"{name}.lengths"], ts=g[f"{name}.ts"],
                x=g[f"{name}.x"], x2=g[f"{name}.x2"], delta=(g[f"{name}.delta0"], g[f"{name}.delta1"]),
                y0=g[f"{name}.y0"], y_inc=g[f"{name}.y_inc"], cache0=cache0, cache1=cache1)
