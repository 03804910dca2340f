import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden" / "golden_v1.json"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(GOLDEN) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_null_models(golden):
    """The reference's own null models (models/null.py run in the build container), in the mirror's layout."""
    from flemingo_b200 import objects
    models = {t: {int(L): {"frequencies": np.array(v["frequencies"], dtype=np.float64), "bins": np.array(v["bins"], dtype=np.float64)}
                  for L, v in d.items()} for t, d in golden["null_models"].items()}
    objects.set_null_models(models)
    return models


@pytest.fixture(scope="session")
def golden_cases(golden, golden_null_models):
    """(case, mirror organism) pairs: organisms rebuilt from the constructor parameters with OUR host objects."""
    from tests.helpers import build_case_organism
    return [(case, build_case_organism(case)) for case in golden["cases"]]


@pytest.fixture(scope="session", params=["filter+refine", "filter+refine, one row", "filter+refine, int32 only",
                                         "filter+refine, pruned stages", "filter+refine, pruned stages, one row", "exact-only"])
def engine(request):
    """The process-wide engine in its kernel modes -- the integer filter pass as the host plans it (packed 16-bit DPX
    where the organism's value range allows it, int32 otherwise; two quantised rows per pair in shared memory), the same
    with ONE row (in-place DP stages; chosen automatically only when it lets more warps live on an SM, forced here), the
    int32 arithmetic for every organism, every organism through the kernels with pruned DP stages (dp_stage_pruned; chosen
    automatically only for organisms with tight connectors) with two rows and in place, and every pair through the exact
    fp64 kernel.  All of them must reproduce the reference bit for bit."""
    import os
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from flemingo_b200.engine import get_engine
    eng = get_engine()
    eng.exact_only = request.param == "exact-only"
    if request.param.endswith("one row"):
        os.environ["FLEMINGO_FAST_ROWS"] = "1"      # read by fl_place_batch at every call
    if request.param.endswith("int32 only"):
        os.environ["FLEMINGO_FAST_BITS"] = "32"
    if "pruned stages" in request.param:
        os.environ["FLEMINGO_PRUNE"] = "2"
    yield eng
    os.environ.pop("FLEMINGO_PRUNE", None)
    os.environ.pop("FLEMINGO_FAST_ROWS", None)
    os.environ.pop("FLEMINGO_FAST_BITS", None)
    eng.exact_only = False
