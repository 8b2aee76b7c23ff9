"""Input pipeline: ml_function_b200.data_prepare against the reference's own
kon/utils/data_prepare.py (executed from /root/reference into tests/golden/ref_layers.npz, case
`data_prepare`) and against the tf.data semantics DP:335-337 asks for."""
import numpy as np
import pandas as pd
import pytest
import torch

from helpers import ref_case
from ml_function_b200 import data_prepare as DPm


def _frames():
    c = ref_case("layers", "data_prepare")
    raw = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ref_layers.npz"))
    s = raw["data_prepare/in/sparse"]
    sdf = pd.DataFrame({"C14": s[:, 0], "C15": s[:, 1], "C16": s[:, 2]}).replace("<nan>", np.nan)
    # column C15 held python ints (object dtype) in the reference run: LabelEncoder sees their str() form
    sdf["C15"] = [np.nan if (isinstance(v, float) and np.isnan(v)) else int(v) for v in sdf["C15"]]
    sdf["C15"] = sdf["C15"].astype(object)
    ddf = pd.DataFrame(raw["data_prepare/in/dense"], columns=["I1", "I2", "I3"])
    return c, raw, sdf, ddf


def test_feature_encoders_match_the_reference_data_prepare():
    c, raw, sdf, ddf = _frames()
    dp = DPm.data_prepare(batch_size=16, device="cpu")
    enc, sinfo = dp.sparse_fea_deal(sdf.copy())
    assert np.array_equal(enc.to_numpy(), raw["data_prepare/out/ids"])                      # LabelEncoder ids, exact
    assert [i.word_size for i in sinfo] == raw["data_prepare/out/word_size"].tolist()
    assert [i.cross_unit for i in sinfo] == raw["data_prepare/out/cross_unit"].tolist()     # embed_dim=8 default
    assert [i.emb_reg for i in sinfo] == raw["data_prepare/out/emb_reg"].tolist()           # 1e-8 default
    assert list(sinfo[0]._fields) == raw["data_prepare/out/fields"].tolist()                # the sparseFea namedtuple
    assert sinfo[0].input_length == 1 and sinfo[0].mask_zero is False and sinfo[0].batch_size == 16
    den, dinfo = dp.dense_fea_deal(ddf.copy())
    assert np.array_equal(den.to_numpy(), raw["data_prepare/out/dense"])                    # MinMaxScaler, bit for bit
    assert [d.fea_name for d in dinfo] == ["I1", "I2", "I3"]


def test_shuffle_buffer_is_tf_datas_local_shuffle():
    rng = np.random.default_rng(5)
    n, S = 20000, 2048
    o = DPm.shuffle_buffer_order(n, S, rng)
    assert sorted(o.tolist()) == list(range(n))                      # a permutation: every element exactly once
    pos = np.empty(n, dtype=np.int64)
    pos[o] = np.arange(n)
    assert (pos >= np.arange(n) - S + 1).all()                       # element i cannot leave before it entered the buffer
    assert (np.arange(n) - pos).max() < S
    assert np.abs(pos - np.arange(n)).mean() > S / 4                 # and it IS shuffled
    # short inputs: the whole dataset fits the buffer -> a full permutation
    o2 = DPm.shuffle_buffer_order(100, S, rng)
    assert sorted(o2.tolist()) == list(range(100))


def test_pipeline_shuffle_repeat_batch_prefetch():
    n, F = 1000, 5
    ids = np.arange(n * F, dtype=np.int32).reshape(n, F)
    dense = np.arange(n, dtype=np.float32).reshape(n, 1).repeat(3, 1)
    labels = np.stack([np.arange(n) % 2 == 0, np.arange(n) % 2 == 1], 1).astype(np.float32)
    dp = DPm.data_prepare(batch_size=64, device="cpu")
    pipe = dp.data_pipeline(((ids, dense), labels))
    batches = list(pipe)
    assert len(batches) == len(pipe) == (2 * n + 63) // 64           # repeat(2), no drop_remainder
    assert all(b[1].shape == (64, F) for b in batches[:-1]) and batches[-1][1].shape[0] == 2 * n - 64 * (len(batches) - 1)
    rows = torch.cat([b[1][:, 0] for b in batches]) // F
    first, second = rows[:n], rows[n:]
    assert sorted(first.tolist()) == list(range(n)) and sorted(second.tolist()) == list(range(n))   # two full epochs
    assert not torch.equal(first, second)                            # reshuffled each repeat
    for d, i, y in batches:                                          # features and labels stay aligned
        r = i[:, 0] // F
        assert torch.equal(d[:, 0], r.float()) and torch.equal(y[:, 0], (r % 2 == 0).float())
        assert i.dtype == torch.int32 and d.dtype == torch.float32


def test_batch_size_none_fails_like_the_reference():
    with pytest.raises(ValueError, match="batch"):
        DPm.data_prepare(batch_size=None, device="cpu").data_pipeline(((np.zeros((4, 2), np.int32), None), np.zeros((4, 2))))


def test_extract_train_test_keeps_features_and_labels_aligned():
    n = 200
    sdf = pd.DataFrame({"a": np.arange(n) % 7, "b": np.arange(n) % 3})
    ddf = pd.DataFrame({"x": np.arange(n, dtype=float) / n})
    target = pd.Series((np.arange(n) % 7 == 0).astype(int))
    dp = DPm.data_prepare(batch_size=32, device="cpu")
    tr, te = dp.extract_train_test(list(range(150)), list(range(150, 200)), target, sparseDf=sdf, denseDf=ddf)
    for d, i, y in tr:
        assert y.shape[1] == 2 and torch.equal(y[:, 1], (i[:, 0] == 0).float())      # to_categorical, aligned with ids
        assert d.shape[0] == i.shape[0] == 32                                        # static batches (DP:390-404)
