"""Streaming writer for ``reps_*.pkl`` / ``query_reps.pkl``.

The reference (dpr_scale/task/dpr_eval_task.py:40-49) keeps every batch on the host, ``torch.cat``s them (the shard is
then in RAM twice: 2 x 8 GB per rank for the 21 M-passage corpus) and ``pickle.dump(tensor, f, protocol=4)``s the result.
This writer produces a file that ``pickle.load`` turns into the SAME object - one contiguous fp32 CPU ``torch.Tensor``
[N, d], which is what run_retrieval_pytorch.py:181 (``torch.tensor(pickle.load(f))``) reads - while holding only the
batches still in flight: rows are appended to the pickle's payload as they arrive.

File layout (what ``pickle.dump`` of a CPU tensor emits, written by hand so that the payload can be streamed):

    PROTO 4
    GLOBAL torch._utils._rebuild_tensor_v2, MARK
      GLOBAL torch.storage._load_from_bytes, BINBYTES8 <len> <blob> -> TUPLE1, REDUCE        # the storage
      0, (N, d), (d, 1), False, OrderedDict()                                               # offset, size, stride, ...
    TUPLE, REDUCE, STOP

``blob`` is torch's legacy (non-zip) serialisation of one FloatStorage: three small pickles (magic, protocol version,
system info), the storage record via a persistent id, the list of storage keys, the element count as int64 and then the
raw little-endian fp32 rows.  Every size field is written with a fixed width (LONG1 with 8 data bytes / int64), so the
header can be emitted before N is known and patched in ``close()``.
"""
import collections
import io
import pickle
import struct

import torch

_MAGIC = 0x1950A86A20F9469CFC6C          # torch.serialization.MAGIC_NUMBER
_PROTOCOL_VERSION = 1001                 # torch.serialization.PROTOCOL_VERSION
_KEY = "0"                               # storage key: any string unique within the file


def _long8(v):
    """pickle LONG1 opcode with exactly 8 payload bytes (fixed width whatever the value)."""
    return b"\x8a\x08" + struct.pack("<q", int(v))


def _body(obj):
    """protocol-4 opcodes that push `obj`, without the PROTO / FRAME prefix and the STOP suffix."""
    b = pickle.dumps(obj, protocol=4)
    assert b[:2] == b"\x80\x04" and b[-1:] == b"."
    b = b[2:-1]
    if b[:1] == b"\x95":                 # FRAME + 8-byte length
        b = b[9:]
    return b


class _StoragePickler(pickle.Pickler):
    """Emits the persistent-id record torch's legacy loader expects for one CPU FloatStorage."""

    def __init__(self, f, numel_field):
        super().__init__(f, protocol=2)
        self.numel_field = numel_field

    def persistent_id(self, obj):
        if obj is _StoragePickler:
            return ("storage", torch.FloatStorage, _KEY, "cpu", self.numel_field, None)
        return None


def _blob_header(numel):
    f = io.BytesIO()
    pickle.dump(_MAGIC, f, protocol=2)
    pickle.dump(_PROTOCOL_VERSION, f, protocol=2)
    pickle.dump({"protocol_version": _PROTOCOL_VERSION, "little_endian": True,
                 "type_sizes": {"short": 2, "int": 4, "long": 4}}, f, protocol=2)
    # storage record: build it with numel = 0x0102030405060708 pickled as LONG1(8) and patch the value afterwards
    probe = 0x0102030405060708
    g = io.BytesIO()
    _StoragePickler(g, probe).dump(_StoragePickler)
    rec = g.getvalue()
    enc = b"\x8a\x08" + struct.pack("<q", probe)
    assert rec.count(enc) == 1, "unexpected integer encoding in the storage record"
    numel_off_in_rec = rec.index(enc) + 2
    rec_off = f.tell()
    f.write(rec.replace(enc, _long8(numel)))
    pickle.dump([_KEY], f, protocol=2)
    count_off = f.tell()
    f.write(struct.pack("<q", int(numel)))
    return f.getvalue(), rec_off + numel_off_in_rec, count_off


class StreamingTensorPickle:
    """``with StreamingTensorPickle(path, dim) as w: w.append(rows) ...`` -> a pickle of ONE fp32 CPU tensor [N, dim]."""

    def __init__(self, path, dim):
        self.path, self.dim, self.rows = path, int(dim), 0
        self.f = open(path, "wb")
        # GLOBAL lookups are written as SHORT_BINUNICODE x2 + STACK_GLOBAL, exactly like the standard pickler
        self.f.write(b"\x80\x04")
        self.f.write(_body("torch._utils") + _body("_rebuild_tensor_v2") + b"\x93\x94" + b"(")
        self.f.write(_body("torch.storage") + _body("_load_from_bytes") + b"\x93\x94")
        self.f.write(b"\x8e")                          # BINBYTES8
        self._len_off = self.f.tell()
        self.f.write(struct.pack("<Q", 0))
        blob_head, self._numel_off, self._count_off = _blob_header(0)
        self._blob_off = self.f.tell()
        self._blob_head_len = len(blob_head)
        self.f.write(blob_head)

    def append(self, rows):
        """rows: fp32 CPU tensor [b, dim] (contiguous); written straight from its buffer."""
        assert rows.dtype == torch.float32 and rows.device.type == "cpu" and rows.dim() == 2 and rows.shape[1] == self.dim
        rows = rows.contiguous()
        self.f.write(memoryview(rows.numpy()).cast("B"))
        self.rows += rows.shape[0]

    def close(self):
        if self.f is None:
            return self.path
        numel = self.rows * self.dim
        f = self.f
        # after the payload: MEMOIZE, TUPLE1, MEMOIZE, REDUCE, MEMOIZE, then offset / size / stride / requires_grad / hooks
        f.write(b"\x94\x85\x94R\x94")
        f.write(b"K\x00")                                             # storage offset 0
        f.write(_long8(self.rows) + _long8(self.dim) + b"\x86\x94")   # size  (N, d)
        f.write(_long8(self.dim) + b"K\x01" + b"\x86\x94")            # stride (d, 1)
        f.write(b"\x89")                                              # requires_grad = False
        f.write(_body(collections.OrderedDict()))                     # backward hooks
        f.write(b"t\x94R\x94.")
        f.seek(self._len_off)
        f.write(struct.pack("<Q", self._blob_head_len + numel * 4))
        f.seek(self._blob_off + self._numel_off)
        f.write(struct.pack("<q", numel))
        f.seek(self._blob_off + self._count_off)
        f.write(struct.pack("<q", numel))
        f.close()
        self.f = None
        return self.path

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
