"""CPU: the analytic op counter reproduces the reference's notebook known answers."""
import json
import os

from wavelet_monodepth_b200 import opcount as oc
from wavelet_monodepth_b200 import synth

from helpers import GOLDEN, golden_names, load_golden

KAT = json.load(open(os.path.join(GOLDEN, "known_answers.json")))


def kitti_total(enc, height, width, counts=None):
    dec = [16, 32, 64, 128, 256]
    per_level = {}
    for i in (4, 3, 2, 1):
        h, w = height >> (i + 1), width >> (i + 1)
        cin0 = enc[-1] if i == 4 else dec[i + 1]
        if i == 4:
            per_level[i - 1] = oc.kitti_level_ops(i, h, w, cin0, dec[i], enc[i - 1], False)
        else:
            m2, m4, m5 = counts[i] if counts else (h * w, 4 * h * w, 4 * h * w)
            per_level[i - 1] = oc.kitti_level_ops(i, h, w, cin0, dec[i], enc[i - 1], True, m2, m4, m5)
    return per_level


def test_kitti_r50_known_answer():
    # KITTI/sparsity_test_notebook.ipynb:1345 -> 17.474 GFLOPs
    per = kitti_total(list(synth.RESNET50_CH), 320, 1024)
    assert sum(per.values()) == 17473692295 == KAT["kitti_r50_1024x320_total_ops"]
    for s in range(4):
        assert per[s] == KAT["kitti_r50_1024x320_total_ops_s%d" % s]


def test_nyu_densenet161_known_answer():
    # NYUv2/sparsity_test_notebook.ipynb:1344 -> 33.464 GFLOPs
    f, h, w = 1104, 15, 20
    t = oc.nyu_dense_part_ops(2208, h, w, f, 384)
    t += oc.nyu_sparse_block_ops(2 * h, 2 * w, f // 2 + 192, f // 4, 16 * h * w, 16 * h * w, False)
    t += oc.nyu_sparse_block_ops(4 * h, 4 * w, f // 4 + 96, f // 8, 64 * h * w, 64 * h * w, True)
    assert t == 33463546800 == KAT["nyu_d161_640x480_total_ops"]


def test_counts_from_golden_masks_give_golden_total_ops():
    """Weight-dependent case: active counts taken from the reference's masks reproduce its total_ops."""
    for name in golden_names("kitti_tiny_sparse"):
        want, meta = load_golden(name)
        counts = {i: (int(want["upconv0_mask_%d" % (i - 1)].sum()), int(want["upconv1_mask_%d" % (i - 1)].sum()),
                      int(want["wavelet_mask_%d" % (i - 1)].sum())) for i in (3, 2, 1)}
        per = kitti_total(list(meta["num_ch_enc"]), meta["height"], meta["width"], counts)
        assert sum(per.values()) == int(want["total_ops"]), name
        for s in range(4):
            assert per[s] == int(want["total_ops_%d" % s]), (name, s)
