"""The product-side synthetic input generators (bench workloads) and the oracle's own copies must
produce identical data: the benchmarks never import ``oracle/`` for their GPU arm."""
import numpy as np

from avdn_b200.utils import synthetic as syn
from oracle import model_oracle as mo
from oracle import warp_oracle as wo


def test_generators_match_the_oracle_copies():
    assert np.array_equal(syn.synthetic_tile(seed=3, size=96), wo.synthetic_tile(seed=3, size=96))
    assert np.array_equal(syn.synthetic_tile(seed=3, size=64, smooth=True), wo.synthetic_tile(seed=3, size=64, smooth=True))
    assert np.array_equal(syn.synthetic_attention_tile(seed=4, size=400), wo.synthetic_attention_tile(seed=4, size=400))
    assert np.array_equal(syn.synthetic_pose_corners(64, seed=5), wo.synthetic_pose_corners(64, seed=5))
    assert syn.yolov3_trunk_cfg() == mo.yolov3_trunk_cfg()
