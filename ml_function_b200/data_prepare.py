"""Input side of the hot path: the reference's ``kon/utils/data_prepare.py`` (= DP) semantics, feeding the
kernels packed ``ids [B,F] int32`` / ``dense [B,n_dense] float32`` / labels directly.

What is kept from the reference (same names, arguments and results):

* ``data_prepare(batch_size, use_shuffle, cpu_core)`` with the ``sparseFea`` / ``denseFea`` descriptors (DP:59-60);
* ``sparse_fea_deal`` (DP:85-102): ``fillna('-1')`` -> per-column ``LabelEncoder`` on the string form -> ids,
  ``word_size = nunique``;  ``dense_fea_deal`` (DP:294-301): mode-fill + ``MinMaxScaler(0,1)``;
* ``concat_test_train`` (DP:78-83), ``static_batch`` (DP:390-404), ``input_loc`` (DP:382-388),
  ``extract_train_test`` (DP:339-380: ``to_categorical`` labels, train/test split, pipeline);
* ``data_pipeline`` (DP:335-337): ``shuffle(2048).repeat(2).batch(batch_size).prefetch(2)``;
* ``FeatureInput`` (DP:65-76) -> ``models.InputFeature``.

What is different: the reference declares ids as float32 Keras Inputs and lets TensorFlow run the
pipeline on the host.  Here the encoded dataset is packed ONCE into three arrays (ids int32 ``[N,F]``, dense float32
``[N,n_dense]``, labels) that live either in HBM (``resident="device"``: a 45 M-row Criteo day is 7 GB of the 180 GB) or
in pinned host memory (``resident="host"``); a batch is one row gather of each array in the shuffled order, the
host->device copies run on a side stream, and ``prefetch(2)`` keeps two batches in flight.  ids stay integers end to end
(float32 cannot address more than 2^24 rows -- BASELINE config 5 has 10^8-row tables).
"""
from __future__ import annotations

import multiprocessing as mp
from collections import namedtuple
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .layers import denseFea, sparseFea


def shuffle_buffer_order(n: int, buffer_size: int, rng: np.random.Generator) -> np.ndarray:
    """Output order of ``tf.data.Dataset.shuffle(buffer_size)`` over ``n`` elements: a buffer of the next
    ``buffer_size`` elements is kept; every output is a uniformly random slot of the buffer, which is then refilled
    with the next input element.  Hence element i can never be emitted before output position ``i - buffer_size + 1``:
    a *local* shuffle, unlike a full permutation.  (Vectorised per output element over a pre-drawn slot stream.)"""
    if n <= 0:
        return np.empty(0, dtype=np.int64)
    bs = int(min(buffer_size, n))
    n_fill = n - bs                                     # outputs during which the buffer stays full
    out = np.empty(n, dtype=np.int64)
    last = np.arange(bs, dtype=np.int64)                # element sitting in each slot when the input runs dry
    if n_fill > 0:
        # Output i takes the element in slot s_i and refills the slot with input element bs + i.  So the element
        # emitted at time i is the one placed at the PREVIOUS time j < i slot s_i was drawn (element bs + j), or the
        # slot's initial element s_i: a grouped "previous occurrence", vectorised with one stable argsort.
        slots = rng.integers(0, bs, size=n_fill)
        by_slot = np.argsort(slots, kind="stable")      # times grouped by slot, ascending inside a group
        ss = slots[by_slot]
        first = np.ones(n_fill, dtype=bool)
        first[1:] = ss[1:] != ss[:-1]
        prev_time = np.empty(n_fill, dtype=np.int64)
        prev_time[1:] = by_slot[:-1]
        val = np.where(first, ss, bs + prev_time)
        out[by_slot] = val
        is_last = np.ones(n_fill, dtype=bool)
        is_last[:-1] = ss[1:] != ss[:-1]
        last[ss[is_last]] = bs + by_slot[is_last]
    out[n_fill:] = last[rng.permutation(bs)]            # draining a buffer by uniform draws = a uniform permutation of it
    return out


class Pipeline:
    """``Dataset.from_tensor_slices((features, labels)).shuffle(S).repeat(R).batch(B).prefetch(P)`` (DP:335-337) over
    packed arrays.  Iterating yields ``(dense [b,n_dense] f32, ids [b,F] i32, labels)`` on ``device``; the last batch of
    the stream may be short (``batch`` without ``drop_remainder``); each repeat reshuffles."""

    def __init__(self, ids, dense, labels, batch_size: Optional[int], device="cuda", shuffle_buffer: int = 2048,
                 repeat: int = 2, prefetch: int = 2, resident: str = "device", seed: int = 2020):
        if batch_size is None:
            raise ValueError("data_prepare(batch_size=None): tf.data's .batch(None) fails in the reference too "
                             "(DP:337); give a batch size")
        self.n = int(ids.shape[0])
        self.batch_size, self.shuffle_buffer, self.repeat, self.prefetch = int(batch_size), int(shuffle_buffer), int(repeat), int(prefetch)
        self.device = torch.device(device)
        self.resident = resident
        self.seed = seed
        t_ids = torch.as_tensor(np.ascontiguousarray(ids)).to(torch.int32)
        t_dense = torch.as_tensor(np.ascontiguousarray(dense), dtype=torch.float32) if dense is not None else torch.zeros((self.n, 0))
        t_lab = torch.as_tensor(np.ascontiguousarray(labels), dtype=torch.float32)
        if resident == "device" and self.device.type == "cuda":
            self.arrays = tuple(t.to(self.device) for t in (t_dense, t_ids, t_lab))
        else:
            pin = self.device.type == "cuda"
            self.arrays = tuple(t.pin_memory() if pin else t for t in (t_dense, t_ids, t_lab))
        self._side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def __len__(self):
        total = self.n * self.repeat
        return (total + self.batch_size - 1) // self.batch_size

    def order(self) -> np.ndarray:
        rng = np.random.default_rng(self.seed)
        return np.concatenate([shuffle_buffer_order(self.n, self.shuffle_buffer, rng) for _ in range(self.repeat)])

    def _make(self, idx: np.ndarray):
        """One batch: a row gather of each packed array (on the device for a resident dataset, else on the host into
        pinned staging followed by an async copy on the side stream)."""
        if self.arrays[0].device.type == "cuda":
            it = torch.as_tensor(idx, device=self.device)
            return tuple(a.index_select(0, it) for a in self.arrays), None
        it = torch.as_tensor(idx)
        host = tuple(a.index_select(0, it) for a in self.arrays)
        if self._side is None:
            return host, None
        host = tuple(h.pin_memory() for h in host)
        with torch.cuda.stream(self._side):
            dev = tuple(h.to(self.device, non_blocking=True) for h in host)
            ev = torch.cuda.Event()
            ev.record(self._side)
        return dev, (ev, host)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        order = self.order()
        nb = (order.shape[0] + self.batch_size - 1) // self.batch_size
        queue: List = []
        nxt = 0
        while nxt < nb or queue:
            while nxt < nb and len(queue) < max(self.prefetch, 1):
                queue.append(self._make(order[nxt * self.batch_size:(nxt + 1) * self.batch_size]))
                nxt += 1
            batch, pending = queue.pop(0)
            if pending is not None:
                torch.cuda.current_stream(self.device).wait_event(pending[0])
                for t in batch:
                    t.record_stream(torch.cuda.current_stream(self.device))
            yield batch


class data_prepare(object):
    """DP:56-64."""

    def __init__(self, batch_size=None, use_shuffle=True, cpu_core=None, device="cuda"):
        self.sparseFea = sparseFea
        self.denseFea = denseFea
        self.batch_size = batch_size
        self.use_shuffle = use_shuffle
        self.cpu_core = mp.cpu_count() if cpu_core is None else cpu_core
        self.device = device

    # ---- feature specs / encoders ---------------------------------------------------------------------
    def FeatureInput(self, sparseInfo: list = None, denseInfo: list = None, seqInfo=None, useLinear: bool = False,
                     useAddLinear: bool = False, useFlattenLinear: bool = False, useFlattenSparse: bool = False):
        from .models import FeatureInput as _FI
        return _FI(sparseInfo, denseInfo, seqInfo, useLinear, useAddLinear, useFlattenLinear, useFlattenSparse,
                   device=self.device)

    def concat_test_train(self, train_df, test_df):
        import pandas as pd
        train_idx = train_df.index.tolist()
        test_idx = list(np.array(test_df.index) + train_idx[-1] + 1)
        df = pd.concat([train_df, test_df], ignore_index=True)
        return df, (train_idx, test_idx)

    def sparse_fea_deal(self, sparseDf, embed_dim=8, linear_dim=1, pre_weight=None, emb_reg=None):
        """DP:85-102.  LabelEncoder on the string form = rank of the value among the sorted unique strings."""
        if not pre_weight:
            pre_weight = [None] * sparseDf.shape[1]
        if not emb_reg:
            emb_reg = [1e-8] * sparseDf.shape[1]
        sparseDf = sparseDf.fillna('-1')
        enc = {}
        for fea in sparseDf:
            col = sparseDf[fea].astype('str').to_numpy()
            uniq, inv = np.unique(col, return_inverse=True)          # sorted unique strings, like LabelEncoder.fit
            enc[fea] = inv.astype(np.int64)
        import pandas as pd
        sparseDf = pd.DataFrame(enc, index=sparseDf.index)
        sparseInfo = [self.sparseFea(
            fea_name=fea, input_dim=sparseDf[fea].shape[0], cross_unit=embed_dim, linear_unit=linear_dim,
            word_size=int(sparseDf[fea].nunique()), pre_weight=weight_, input_length=1, is_trainable=True,
            mask_zero=False, sample_num=None, batch_size=self.batch_size, emb_reg=reg
        ) for fea, weight_, reg in zip(sparseDf, pre_weight, emb_reg)]
        return sparseDf, sparseInfo

    def dense_fea_deal(self, denseDf, is_fillna=True):
        """DP:294-301: mode-fill, then ``MinMaxScaler(feature_range=(0,1))`` per column
        (``(x - min) / (max - min)``, constant columns map to 0)."""
        import pandas as pd
        if is_fillna:
            denseDf = pd.DataFrame({fea: denseDf[fea].fillna(denseDf[fea].mode()[0]) for fea in denseDf})
        x = denseDf.to_numpy(dtype=np.float64)
        lo, hi = np.nanmin(x, axis=0), np.nanmax(x, axis=0)
        rng = hi - lo
        rng[rng == 0.0] = 1.0                                        # sklearn's _handle_zeros_in_scale
        scale = 1.0 / rng
        scaled = x * scale + (0.0 - lo * scale)                       # sklearn: X * scale_ + min_
        denseDf = pd.DataFrame(scaled, columns=denseDf.columns, index=denseDf.index)
        denseInfo = [self.denseFea(fea, self.batch_size) for fea in denseDf]
        return denseDf, denseInfo

    # ---- batching ---------------------------------------------------------------------------------------
    def input_loc(self, df, use_idx: list):
        if isinstance(df, dict):
            return {key: np.array(df[key])[use_idx] for key in df}
        return df[use_idx]

    def static_batch(self, df):
        """DP:390-404: sample (with replacement, like ``np.random.choice``) a multiple of the batch size."""
        df_num = np.array(df[list(df.keys())[0]]).shape[0] if isinstance(df, dict) else len(df)
        batch_num = (df_num // self.batch_size) * self.batch_size
        need_idx = np.random.choice(list(range(df_num)), size=batch_num)
        if self.use_shuffle:
            np.random.shuffle(need_idx)
        return self.input_loc(df, use_idx=need_idx)

    def pack(self, sparseDf=None, denseDf=None):
        """DataFrames of encoded features -> (ids int32 [N,F], dense float32 [N,n_dense])."""
        ids = np.stack([sparseDf[c].to_numpy() for c in sparseDf], axis=1).astype(np.int32) if sparseDf is not None else None
        dense = np.stack([denseDf[c].to_numpy() for c in denseDf], axis=1).astype(np.float32) if denseDf is not None else None
        return ids, dense

    def data_pipeline(self, dataSet: tuple, resident: str = "device", seed: int = 2020) -> Pipeline:
        """DP:335-337 on packed arrays: ``dataSet = ((ids, dense), labels)``."""
        (ids, dense), labels = dataSet
        return Pipeline(ids, dense, labels, self.batch_size, device=self.device, resident=resident, seed=seed)

    def extract_train_test(self, train_idx, test_idx, targetDf, sparseDf=None, denseDf=None, seqDf=None,
                           use_softmax=True, resident: str = "device"):
        """DP:339-380 (sequence features are outside this build: ``seqDf`` must be None)."""
        if seqDf is not None:
            raise NotImplementedError("sequence features: use SparseEmbed / SeqBaseLayer directly")
        y = np.asarray(targetDf.values.tolist())
        if use_softmax:                                              # tf.keras.utils.to_categorical
            yi = y.astype(np.int64).ravel()
            onehot = np.zeros((yi.shape[0], int(yi.max()) + 1), dtype=np.float32)
            onehot[np.arange(yi.shape[0]), yi] = 1
            y = onehot
        ids, dense = self.pack(sparseDf, denseDf)
        out = []
        for idx in (train_idx, test_idx):
            idx = np.asarray(idx)
            if self.batch_size is not None:
                # DP:370-374 static-batches features and labels with two INDEPENDENT random draws, which
                # mis-aligns them; here one draw indexes both (the one deliberate divergence of this module)
                n_keep = (idx.shape[0] // self.batch_size) * self.batch_size
                pick = np.random.choice(idx.shape[0], size=n_keep)
                if self.use_shuffle:
                    np.random.shuffle(pick)
                idx = idx[pick]
            part = ((ids[idx] if ids is not None else None, dense[idx] if dense is not None else None), y[idx])
            out.append(self.data_pipeline(part, resident=resident))
        return out[0], out[1]
