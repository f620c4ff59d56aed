import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def rel_err(a, b):
    """Norm-wise relative error |a-b| / |b| in float64."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return num / den if den > 0 else num


L2_CASES = ["l2_attr_stopgrad", "l2_attr_stopgrad_gq_only", "l2_attr_first_n", "l2_attr_st_onehot",
            "l2_attr_st_onehot_first_n", "l2_attr_learn_temp", "l2_attr_temp_quarter",
            "l2_noattr_k37_d32", "l2_noattr_k300_d128", "l2_attr_skip_train", "l2_attr_ragged",
            "l2_attr_ragged_b3_s37", "l2_config1_16x200",
            # enc_embs produced by the reference's own CTC encoder (oracle/gen_golden_realistic.py)
            "l2_realistic_a", "l2_realistic_b"]
SEP_CASES = ["sep_attr_stopgrad", "sep_attr_st_onehot", "sep_noattr_k29_d48", "sep_noattr_st_onehot",
             "sep_config1_16x200"]
# cases whose reference module was built with stop_grad=False
ST_ONEHOT = {"l2_attr_st_onehot", "l2_attr_st_onehot_first_n", "sep_attr_st_onehot", "sep_noattr_st_onehot"}


@pytest.fixture
def golden():
    return load_golden
