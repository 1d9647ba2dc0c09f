"""Scheme-driven episode container with the reference's API (marl/components/episode_batch.py:57-248),
re-laid-out for HBM: every episode is ONE 128-byte aligned *record* that holds all scheme keys, and the per-key
tensors the reference exposes (``batch["obs"]`` -> ``[B, T+1, N, OBS]``) are strided views into the record array.

Why: ReplayBuffer.sample / insert_episode_batch then move whole records with one bulk-copy kernel launch
(``mal_record_copy``: cp.async.bulk HBM -> smem -> HBM) instead of one advanced-index copy per key, and the learner
kernels read each field through (pointer, batch stride, time stride).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace as SN

import numpy as np
import torch as th

from .. import _native as nat

RECORD_ALIGN = 128


def _check_safe_view(v, dest, key):
    # episode_batch.py:6-13 -- trailing dims must match unless the destination dim is 1
    idx = len(v.shape) - 1
    for s in dest.shape[::-1]:
        if v.shape[idx] != s:
            if s != 1:
                raise ValueError("Unsafe reshape of {} to {} at Key: {}".format(v.shape, dest.shape, key))
        else:
            idx -= 1


def _get_num_items(indexing_item, max_size):
    # episode_batch.py:16-22 (tensor indices yield None there; we count them, which is what callers expect)
    if isinstance(indexing_item, (list, np.ndarray)):
        return len(indexing_item)
    if isinstance(indexing_item, th.Tensor):
        return int(indexing_item.numel())
    if isinstance(indexing_item, slice):
        r = indexing_item.indices(max_size)
        return 1 + (r[1] - r[0] - 1) // r[2]


def _new_data_sn():
    d = SN()
    d.transition_data = {}
    d.episode_data = {}
    return d


def _is_index_array(item):
    return isinstance(item, (list, np.ndarray)) or (isinstance(item, th.Tensor) and item.dtype == th.long)


def _parse_slices(items):
    # episode_batch.py:33-54
    if isinstance(items, (slice, int)) or _is_index_array(items):
        items = (items, slice(None))
    if isinstance(items[1], list):
        raise IndexError("Indexing across Time must be contiguous")
    parsed = []
    for item in items:
        parsed.append(slice(item, item + 1) if isinstance(item, int) else item)
    return parsed


def _prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class RecordLayout:
    """Byte layout of one episode record: per key an aligned section [T, *shape] (or [*shape] if episode_const)."""

    def __init__(self, fields, max_seq_length):
        self.max_seq_length = max_seq_length
        self.fields = {}          # key -> (offset_bytes, shape, dtype, episode_const)
        off = 0
        for key, (shape, dtype, const) in fields.items():
            itemsize = th.empty((), dtype=dtype).element_size()
            n = _prod(shape) * (1 if const else max_seq_length) * itemsize
            self.fields[key] = (off, tuple(shape), dtype, const)
            off += -(-n // RECORD_ALIGN) * RECORD_ALIGN
        self.record_bytes = max(off, RECORD_ALIGN)

    def same_as(self, other):
        return (other is not None and self.record_bytes == other.record_bytes
                and self.max_seq_length == other.max_seq_length and self.fields == other.fields)

    def views(self, storage, batch_size):
        """Per-key strided views [B, (T,) *shape] into the uint8 record array `storage`."""
        out = {}
        for key, (off, shape, dtype, const) in self.fields.items():
            itemsize = th.empty((), dtype=dtype).element_size()
            typed = storage.view(dtype)
            sizes, strides, acc = [], [], 1
            for s in reversed(shape):
                sizes.append(s)
                strides.append(acc)
                acc *= s
            if not const:
                sizes.append(self.max_seq_length)
                strides.append(acc)
            sizes.append(batch_size)
            strides.append(self.record_bytes // itemsize)
            out[key] = typed.as_strided(tuple(reversed(sizes)), tuple(reversed(strides)), off // itemsize)
        return out


class EpisodeBatch:
    def __init__(self, scheme, groups, batch_size, max_seq_length, data=None, preprocess=None, device="cpu"):
        self.scheme = scheme.copy()
        self.groups = groups
        self.batch_size = batch_size
        self.max_seq_length = max_seq_length
        self.preprocess = {} if preprocess is None else preprocess
        self.device = device
        self._layout = None       # RecordLayout when this object owns a packed record array
        self._storage = None      # uint8 [batch_size * record_bytes]
        if data is not None:
            self.data = data
        else:
            self.data = _new_data_sn()
            self._setup_data(self.scheme, self.groups, batch_size, max_seq_length, self.preprocess)

    # ------------------------------------------------------------------ construction
    def _setup_data(self, scheme, groups, batch_size, max_seq_length, preprocess):
        # scheme derivation follows episode_batch.py:89-143
        if preprocess is not None:
            for k in preprocess:
                assert k in scheme
                new_k, transforms = preprocess[k][0], preprocess[k][1]
                vshape, dtype = self.scheme[k]["vshape"], self.scheme[k]["dtype"]
                for transform in transforms:
                    vshape, dtype = transform.infer_output_info(vshape, dtype)
                self.scheme[new_k] = {"vshape": vshape, "dtype": dtype}
                if "group" in self.scheme[k]:
                    self.scheme[new_k]["group"] = self.scheme[k]["group"]
                if "episode_const" in self.scheme[k]:
                    self.scheme[new_k]["episode_const"] = self.scheme[k]["episode_const"]
        assert "filled" not in scheme, '"filled" is a reserved key for masking.'
        scheme.update({"filled": {"vshape": (1,), "dtype": th.long}})

        fields = {}
        for field_key, info in scheme.items():
            assert "vshape" in info, "Scheme must define vshape for {}".format(field_key)
            vshape = info["vshape"]
            if isinstance(vshape, int):
                vshape = (vshape,)
            group = info.get("group", None)
            if group:
                assert group in groups, "Group {} must have its number of members defined in _groups_".format(group)
                shape = (groups[group], *vshape)
            else:
                shape = tuple(vshape)
            fields[field_key] = (shape, info.get("dtype", th.float32), info.get("episode_const", False))

        if self._layout is None:
            self._layout = RecordLayout(fields, max_seq_length)
            self._storage = th.zeros(batch_size * self._layout.record_bytes, dtype=th.uint8, device=self.device)
            self._bind_views()
        else:  # extend(): extra keys get their own (unpacked) tensors; record copies then fall back per key
            for key, (shape, dtype, const) in fields.items():
                if key in self.data.transition_data or key in self.data.episode_data:
                    continue
                if const:
                    self.data.episode_data[key] = th.zeros((batch_size, *shape), dtype=dtype, device=self.device)
                else:
                    self.data.transition_data[key] = th.zeros((batch_size, max_seq_length, *shape), dtype=dtype,
                                                              device=self.device)
            self._layout = None

    def _bind_views(self):
        views = self._layout.views(self._storage, self.batch_size)
        for key, (_, _, _, const) in self._layout.fields.items():
            (self.data.episode_data if const else self.data.transition_data)[key] = views[key]

    def extend(self, scheme, groups=None):
        self._setup_data(scheme, self.groups if groups is None else groups, self.batch_size, self.max_seq_length, None)

    def to(self, device):
        if self._layout is not None:
            self._storage = self._storage.to(device)
            self.device = device
            self._bind_views()
            return
        for k, v in self.data.transition_data.items():
            self.data.transition_data[k] = v.to(device)
        for k, v in self.data.episode_data.items():
            self.data.episode_data[k] = v.to(device)
        self.device = device

    # ------------------------------------------------------------------ writes
    def update(self, data, bs=slice(None), ts=slice(None), mark_filled=True):
        """episode_batch.py:157-195: cast to the scheme dtype, view into the slice, re-apply preprocess transforms."""
        slices = tuple(_parse_slices((bs, ts)))
        for key, value in data.items():
            if key in self.data.transition_data:
                target = self.data.transition_data
                if mark_filled:
                    target["filled"][slices] = 1
                    mark_filled = False
                _slices = slices
            elif key in self.data.episode_data:
                target = self.data.episode_data
                _slices = slices[0]
            else:
                raise KeyError("{} not found in transition or episode data".format(key))
            dtype = self.scheme[key].get("dtype", th.float32)
            if isinstance(value, th.Tensor):
                value = value.to(dtype).to(device=self.device)
            else:
                value = th.tensor(np.asarray(value), dtype=dtype, device=self.device)
            dest = target[key][_slices]
            _check_safe_view(value, dest, key)
            target[key][_slices] = value.view_as(dest)
            if key in self.preprocess:
                new_k = self.preprocess[key][0]
                value = target[key][_slices]
                for transform in self.preprocess[key][1]:
                    value = transform.transform(value)
                dest = target[new_k][_slices]
                _check_safe_view(value, dest, key)
                target[new_k][_slices] = value.view_as(dest)

    # ------------------------------------------------------------------ reads
    def __getitem__(self, item):
        if isinstance(item, str):
            if item in self.data.episode_data:
                return self.data.episode_data[item]
            if item in self.data.transition_data:
                return self.data.transition_data[item]
            raise ValueError
        if isinstance(item, tuple) and all(isinstance(it, str) for it in item):
            new_data = _new_data_sn()
            for key in item:
                if key in self.data.transition_data:
                    new_data.transition_data[key] = self.data.transition_data[key]
                elif key in self.data.episode_data:
                    new_data.episode_data[key] = self.data.episode_data[key]
                else:
                    raise KeyError("Unrecognised key {}".format(key))
            new_scheme = {key: self.scheme[key] for key in item}
            new_groups = {self.scheme[key]["group"]: self.groups[self.scheme[key]["group"]]
                          for key in item if "group" in self.scheme[key]}
            return EpisodeBatch(new_scheme, new_groups, self.batch_size, self.max_seq_length, data=new_data,
                                device=self.device)
        item = _parse_slices(item)
        ret_bs = _get_num_items(item[0], self.batch_size)
        ret_max_t = _get_num_items(item[1], self.max_seq_length)
        if _is_index_array(item[0]) and self._layout is not None:
            # index arrays copy (episode_batch.py:226-238): gather whole records, then view the time slice
            gathered = self._gather_records(item[0])
            if item[1] == slice(None):
                return gathered
            return gathered[slice(None), item[1]]
        new_data = _new_data_sn()
        idx = tuple(item)
        for k, v in self.data.transition_data.items():
            new_data.transition_data[k] = v[idx]
        for k, v in self.data.episode_data.items():
            new_data.episode_data[k] = v[idx[0]]
        ret = EpisodeBatch(self.scheme, self.groups, ret_bs, ret_max_t, data=new_data, device=self.device)
        if (self._layout is not None and isinstance(item[0], slice) and isinstance(item[1], slice)
                and item[0].indices(self.batch_size)[2] == 1
                and item[1].indices(self.max_seq_length) == (0, self.max_seq_length, 1)):
            # contiguous run of whole records: keeps the bulk-copy insert path usable for wrap-around splits
            ret._parent_records = (self._layout, self._storage, item[0].indices(self.batch_size)[0])
        return ret

    def _gather_records(self, ids):
        """New packed batch holding records `ids` (one bulk-copy launch on CUDA; host memory: one index_select)."""
        if isinstance(ids, th.Tensor):
            ids_t = ids.to(dtype=th.long)          # device index tensors are taken as in-range and non-negative
        else:
            arr = np.asarray(ids, dtype=np.int64).reshape(-1)
            arr = np.where(arr < 0, arr + self.batch_size, arr)
            if arr.size and (arr.min() < 0 or arr.max() >= self.batch_size):
                raise IndexError("episode index out of range")
            ids_t = th.from_numpy(arr)
        n = int(ids_t.numel())
        out = EpisodeBatch.__new__(EpisodeBatch)
        out.scheme, out.groups, out.batch_size, out.max_seq_length = self.scheme, self.groups, n, self.max_seq_length
        out.preprocess, out.device = self.preprocess, self.device
        out.data = _new_data_sn()
        out._layout = self._layout
        rb = self._layout.record_bytes
        dev = self._storage.device
        if dev.type == "cuda":
            out._storage = th.empty(n * rb, dtype=th.uint8, device=dev)
            ids_d = ids_t.to(dev, non_blocking=True)
            if n:
                with th.cuda.device(dev):
                    nat.check(nat.lib().mal_record_copy(nat.ptr(out._storage), rb, None, nat.ptr(self._storage), rb,
                                                        nat.ptr(ids_d), n, rb, nat.current_stream(dev)),
                              "mal_record_copy")
        else:  # host-resident buffer (buffer_cpu_only=True): plain host gather of whole records
            out._storage = self._storage.view(self.batch_size, rb).index_select(0, ids_t.to(dev)).reshape(-1)
        out._bind_views()
        return out

    # ------------------------------------------------------------------ compact wire form (host <-> device path)
    def _wire_batch_struct(self):
        """(mal_batch_t of this packed device batch, wire layout) for the hot-path scheme of ma_experiment.py:99-118."""
        cached = getattr(self, "_wire_cache", None)
        if cached is not None and cached[0] == self._storage.data_ptr():
            return cached[1], cached[2]
        tv = self.data.transition_data
        need = ("state", "obs", "actions", "avail_actions", "reward", "terminated", "actions_onehot", "filled")
        if self._layout is None or set(tv) != set(need) or self.data.episode_data:
            raise nat.MalError("the wire form covers packed batches with exactly the hot-path scheme %s" % (need,))
        obs = nat.require_cuda(tv["obs"], "batch")
        B, TT, N, OBS = obs.shape
        A, S = tv["avail_actions"].shape[-1], tv["state"].shape[-1]
        b = nat.Batch(B, TT, N, A, OBS, S)
        b.obs = nat.field_of(obs, N * OBS)
        b.onehot = nat.field_of(tv["actions_onehot"], N * A)
        b.actions = nat.field_of(tv["actions"], N)
        b.avail = nat.field_of(tv["avail_actions"], N * A)
        b.state = nat.field_of(tv["state"], S)
        b.reward = nat.field_of(tv["reward"], 1)
        b.terminated = nat.field_of(tv["terminated"], 1)
        b.filled = nat.field_of(tv["filled"], 1)
        wl = nat.WireLayout()
        nat.check(nat.lib().mal_wire_layout(TT, N, OBS, S, C.byref(wl)), "mal_wire_layout")
        self._wire_cache = (self._storage.data_ptr(), b, wl)       # the views are fixed once the record array exists
        return b, wl

    def wire_bytes(self):
        """Bytes per episode of the compact wire record."""
        return int(self._wire_batch_struct()[1].record_bytes)

    def to_wire(self, out=None, check=True):
        """This (device-resident, packed) batch as compact wire records: uint8 [batch_size, wire_bytes] on the device.
        The wire form drops what the device can re-derive (`actions_onehot`, `filled` as a bit) and narrows 0/1 flags and
        action indices: 26 % fewer bytes at 5v5.  It is what a host-resident replay buffer
        (`buffer_cpu_only`) should keep in pinned memory and ship per learner step; `load_wire` restores the batch bit for
        bit.  `check=True` raises ValueError (one device sync) if a value does not fit the encoding."""
        b, wl = self._wire_batch_struct()
        dev = self._storage.device
        if out is None:
            out = th.empty(self.batch_size, wl.record_bytes, dtype=th.uint8, device=dev)
        status = th.zeros(1, dtype=th.int32, device=dev) if check else None
        with nat.on_device(dev):
            nat.check(nat.lib().mal_wire_pack(C.byref(b), nat.ptr(out), C.byref(wl), nat.ptr(status), nat.current_stream(dev)),
                      "mal_wire_pack")
        if check and int(status.item()) != 0:
            raise ValueError("batch does not fit the wire encoding (actions in [0, 255], 0/1 avail flags, one-hot = OneHot(actions) on filled steps)")
        return out

    def load_wire(self, wire):
        """Overwrite this packed device batch with the episodes of `wire` (uint8 [batch_size, wire_bytes] on the same device,
        e.g. the H2D copy of pinned host wire records): one unpack launch, bit-exact inverse of `to_wire`."""
        b, wl = self._wire_batch_struct()
        dev = self._storage.device
        if wire.device != dev or wire.dtype != th.uint8 or wire.numel() != self.batch_size * wl.record_bytes or not wire.is_contiguous():
            raise nat.MalError("wire buffer must be a contiguous uint8 [batch_size, %d] tensor on %s" % (wl.record_bytes, dev))
        with nat.on_device(dev):
            nat.check(nat.lib().mal_wire_unpack(C.byref(b), nat.ptr(wire), C.byref(wl), nat.current_stream(dev)), "mal_wire_unpack")
        return self

    def max_t_filled(self):
        """episode_batch.py:240-242; returns a 1-element long tensor like the reference."""
        filled = self.data.transition_data["filled"]
        if filled.is_cuda:
            out = th.empty(1, dtype=th.int32, device=filled.device)
            with th.cuda.device(filled.device):
                nat.check(nat.lib().mal_max_t_filled(nat.ptr(filled), filled.stride(0), filled.stride(1),
                                                     filled.shape[0], filled.shape[1], nat.ptr(out),
                                                     nat.current_stream(filled.device)), "mal_max_t_filled")
            return out.long()
        return th.sum(filled, 1).max(0)[0]

    def __repr__(self):
        return "EpisodeBatch. Batch Size:{} Max_seq_len:{} Keys:{} Groups:{}".format(
            self.batch_size, self.max_seq_length, self.scheme.keys(), self.groups.keys())
