"""CPU: window post-processing and clip collation (host logic either side of the temporal head) against the
reference-pinned oracle (oracle/post_oracle.py) and the golden outputs of the reference's own functions."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_post as MGP
from oracle import post_oracle as PO
from oracle import sais_oracle as O
from sais_b200 import postprocess as PP


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "postprocess.npz")


@pytest.mark.parametrize("seed", MGP.SEEDS)
def test_oracle_and_product_match_reference_goldens(gold, seed):
    reps, protos = MGP.make_inputs(seed)
    P = torch.vstack(list(protos.values()))
    views = [O.prototype_probs(r, P)[0].numpy() for r in reps]  # oracle restatement of calcProbs
    np.testing.assert_allclose(np.stack(views), gold[f"views_{seed}"], rtol=2e-6, atol=1e-7)
    ens_o = PO.ensemble_tta(gold[f"views_{seed}"])
    ens_p = PP.tta_mean(gold[f"views_{seed}"])
    np.testing.assert_allclose(ens_o, gold[f"ens_{seed}"], rtol=1e-6)
    np.testing.assert_allclose(ens_p, gold[f"ens_{seed}"], rtol=1e-6)
    for thr, tag in ((None, "argmax"), (0.515, "thr")):
        for fn in (PO.get_preds, PP.get_preds):
            ent, pred = fn(gold[f"ens_{seed}"], thr)
            np.testing.assert_allclose(ent, gold[f"entropy_{seed}"], rtol=1e-6)
            assert np.array_equal(pred, gold[f"pred_{seed}_{tag}"])
    assert 0 < gold[f"pred_{seed}_thr"].sum() < len(gold[f"pred_{seed}_thr"])  # both classes occur


@pytest.mark.parametrize("k", range(len(MGP.INTERVAL_CASES)))
@pytest.mark.parametrize("seconds", [2, 3])
def test_group_prediction_intervals(gold, k, seconds):
    idx = MGP.INTERVAL_CASES[k]
    for fn in (PO.group_prediction_intervals, PP.group_prediction_intervals):
        s, e = fn(idx, seconds)
        assert list(s) == list(gold[f"int_{k}_{seconds}_s"]) and list(e) == list(gold[f"int_{k}_{seconds}_e"])


def test_group_prediction_intervals_property():
    """every kept index lies in exactly the interval the reference would report, on random index sets"""
    rng = np.random.default_rng(0)
    for _ in range(200):
        idx = sorted(set(rng.integers(0, 60, size=rng.integers(1, 25)).tolist()))
        assert PP.group_prediction_intervals(idx, 3) == PO.group_prediction_intervals(idx, 3)
    assert PP.group_prediction_intervals([], 3) == ([], [])


def test_padding_mask_and_collate_match_reference_semantics():
    lens = [5, 9, 1]
    m = PP.create_padding_mask(lens)
    assert m.shape == (3, 1, 10) and m.dtype == torch.bool
    ref = O.padding_mask(lens, 9)  # oracle restatement of createPaddingMask
    assert torch.equal(m, ref)
    clips = [torch.randn(1, n, 384) for n in lens]
    x, mask, got = PP.pad_collate(clips)
    assert x.shape == (3, 1, 9, 384) and got == lens and torch.equal(mask, m)
    for b, n in enumerate(lens):
        assert torch.equal(x[b, 0, :n], clips[b][0]) and torch.all(x[b, 0, n:] == 0)
    # TTA tuple form: three versions per sample -> three (padded, mask, lens) entries
    tta = [tuple(torch.randn(1, n - o, 384) for o in (0, 3, 6)) for n in (15, 12)]
    xs, ms, ls = PP.pad_collate(tta)
    assert [t.shape for t in xs] == [(2, 1, 15, 384), (2, 1, 12, 384), (2, 1, 9, 384)]
    assert ls == [[15, 12], [12, 9], [9, 6]] and ms[1].shape == (2, 1, 13)


def test_gestures_for_video_filters_and_merges():
    # windows 0-2 confidently class 1, window 3 uncertain (entropy > 0.66), windows 8-9 class 1 again, rest class 0
    p1 = np.array([0.9, 0.85, 0.8, 0.52, 0.1, 0.1, 0.1, 0.1, 0.95, 0.9])
    probs = np.stack([1 - p1, p1], 1)
    starts = np.arange(10) * 15
    ends = starts + 15
    g = PP.gestures_for_video(probs, starts, ends, ["in-view", "out-of-view"], threshold=0.515, entropy_thresh=0.66,
                              seconds=3)
    oov = [(d["StartFrame"], d["EndFrame"]) for d in g if d["Gesture"] == "out-of-view"]
    inv = [(d["StartFrame"], d["EndFrame"]) for d in g if d["Gesture"] == "in-view"]
    assert oov == [(0, 45), (120, 150)] and inv == [(60, 120)]
    ent, _ = PP.get_preds(probs)
    assert ent[3] > 0.66  # the uncertain window was dropped, not merged
    assert PP.frames_to_time(30 * 3725) == (1, 2, 5)


def test_reps_and_labels_round_trip(tmp_path):
    views = tuple(torch.randn(4, 256) for _ in range(3))
    path = PP.save_reps_and_labels(str(tmp_path), "inference", views, videonames=list("abcd"))
    info = torch.load(path)
    assert set(info) == {"reps", "labels", "videonames", "logits"} and len(info["reps"]) == 3
    # exactly what process_inference_results.calcProbs does with the file (:77)
    stacked = torch.stack(info["reps"][1])
    assert torch.equal(stacked, views[1])
    apath = PP.save_attention(str(tmp_path), "inference", [torch.rand(2, 16, 16)])
    assert torch.load(apath)[0].shape == (2, 16, 16)


def test_h5_round_trip(tmp_path):
    """saveH5 / the h5py reads of prepare_dataset.py through the package (h5py when installed, else sais_b200.h5lite)."""
    reps = torch.randn(7, 384)
    labels = ["vidA"] * 3 + ["vidB"] * 4
    path = PP.save_h5(str(tmp_path), "ViT_SelfSupervised_ImageNet", reps, labels)
    assert path.endswith("ViT_SelfSupervised_ImageNet_RepsAndLabels.h5")
    back = PP.load_h5(path)
    assert set(back) == {"vidA", "vidB"} and np.array_equal(back["vidB"], reps[3:].numpy())
    assert back["vidA"].dtype == np.float32 and back["vidA"].shape == (3, 384)
    fpath = PP.save_h5(str(tmp_path), "ViT_SelfSupervised_ImageNet", reps, labels, kind="flow")
    assert fpath.endswith("ViT_SelfSupervised_ImageNet_FlowRepsAndLabels.h5")


_LIBHDF5_FILE = os.path.join(os.path.dirname(__import__("scipy").__file__), "io", "matlab", "tests", "data",
                             "testhdf5_7.4_GLNX86.mat")


@pytest.mark.skipif(not os.path.exists(_LIBHDF5_FILE), reason="SciPy's MATLAB-7.3 (HDF5) test file is not installed")
def test_h5lite_reads_a_libhdf5_file():
    """The reader against a file libhdf5 itself wrote (MATLAB 7.3 = HDF5 behind a 512-byte user block): version-0
    superblock with a base address, symbol-table group, B-tree / SNOD / local heap, version-1 object header with six
    messages, float64 dataset.  SciPy's own test expects exactly linspace(0, 2 pi, 9) in 'testdouble'."""
    from sais_b200 import h5lite

    d = h5lite.describe(_LIBHDF5_FILE)
    assert d["base"] == 512 and d["members"] == ["testdouble"] and (d["leaf_k"], d["int_k"]) == (4, 16)
    got = h5lite.read(_LIBHDF5_FILE)["testdouble"]
    assert got.dtype == np.float64 and got.shape == (9, 1)
    assert np.array_equal(got.ravel(), np.linspace(0, 2 * np.pi, 9))


def test_h5lite_writer_layout_and_round_trip(tmp_path):
    """The writer through the libhdf5-pinned reader: many videos (more than one default symbol-table node holds), names in
    strcmp order, dtypes, empty datasets; message encodings compared byte for byte with what libhdf5 wrote in the file above."""
    import struct
    from sais_b200 import h5lite

    rng = np.random.default_rng(0)
    data = {f"P-{i:03d}_video{'_long_suffix' * (i % 3)}": rng.standard_normal((1 + i % 5, 384)).astype(np.float32)
            for i in range(300)}
    data["zzz_empty"] = np.zeros((0, 384), dtype=np.float32)
    data["Upper"] = np.arange(6, dtype=np.int64).reshape(2, 3)
    data["doubles"] = np.linspace(0, 2 * np.pi, 9).reshape(9, 1)
    path = str(tmp_path / "many.h5")
    h5lite.write(path, data)
    back = h5lite.read(path)
    assert set(back) == set(data)
    for k, v in data.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    d = h5lite.describe(path)
    assert d["members"] == sorted(data, key=lambda s: s.encode()) and d["base"] == 0 and d["root_cache_type"] == 1
    assert d["eof"] == os.path.getsize(path) and d["leaf_k"] * 2 >= len(data)
    with pytest.raises(h5lite.H5LiteError):
        h5lite.write(str(tmp_path / "bad.h5"), {"a/b": np.zeros(3, np.float32)})
    with pytest.raises(h5lite.H5LiteError):
        h5lite.read(__file__)
    if os.path.exists(_LIBHDF5_FILE):
        # same float64 [9,1] dataset as the libhdf5-written file: datatype and dataspace messages must be byte-identical
        theirs = h5lite._Reader(open(_LIBHDF5_FILE, "rb").read())
        ours = h5lite._Reader(open(path, "rb").read())
        t_msgs = dict(theirs._messages(dict(theirs.group_members(theirs.root_entry["ohdr"]))["testdouble"]))
        o_msgs = dict(ours._messages(dict(ours.group_members(ours.root_entry["ohdr"]))["doubles"]))
        assert o_msgs[0x0003] == t_msgs[0x0003] and o_msgs[0x0001] == t_msgs[0x0001] and o_msgs[0x0005] == t_msgs[0x0005]
        addr, size = struct.unpack_from("<QQ", o_msgs[0x0008], 2)
        assert o_msgs[0x0008][:2] == b"\x03\x01" and size == 72 and addr % 8 == 0
