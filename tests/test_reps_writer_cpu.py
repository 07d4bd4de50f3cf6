"""Streaming reps_*.pkl writer (dpr_scale_b200/utils/reps_writer.py): the file must unpickle - with plain ``pickle.load``,
as /root/reference/dpr_scale/run_retrieval_pytorch.py:181 and dpr_eval_task.py:46 imply - to ONE contiguous fp32 CPU
tensor equal to the concatenation of the appended batches, for empty, tiny, ragged and > 2^16-element shards."""
import pickle

import pytest
import torch

from dpr_scale_b200.utils.reps_writer import StreamingTensorPickle


@pytest.mark.parametrize("batches,dim", [([], 8), ([1], 4), ([3, 5, 2], 16), ([128] * 9 + [77], 768), ([70000], 3)])
def test_streamed_pickle_equals_dumped_tensor(tmp_path, batches, dim):
    g = torch.Generator().manual_seed(len(batches) * 31 + dim)
    parts = [torch.randn(b, dim, generator=g) for b in batches]
    path = tmp_path / "reps_0000.pkl"
    with StreamingTensorPickle(str(path), dim) as w:
        for p in parts:
            w.append(p)
    with open(path, "rb") as f:
        got = pickle.load(f)
        assert f.read() == b""                                   # nothing after STOP
    want = torch.cat(parts, 0) if parts else torch.empty(0, dim)
    assert isinstance(got, torch.Tensor) and got.dtype == torch.float32 and got.device.type == "cpu"
    assert got.shape == want.shape and got.is_contiguous() and not got.requires_grad
    assert torch.equal(got, want)
    assert torch.equal(torch.tensor(got), want)                   # the reference's loader: torch.tensor(pickle.load(f))
    # (byte identity with pickle.dump is not a meaningful target: torch embeds a pointer-derived storage key)
    assert torch.equal(pickle.loads(pickle.dumps(got, protocol=4)), want)
    got.add_(1.0)                                                  # the loaded storage is writable, as with pickle.dump


def test_append_rejects_wrong_shapes(tmp_path):
    w = StreamingTensorPickle(str(tmp_path / "x.pkl"), 4)
    with pytest.raises(AssertionError):
        w.append(torch.zeros(2, 5))
    with pytest.raises(AssertionError):
        w.append(torch.zeros(2, 4, dtype=torch.float64))
    w.close()
    assert pickle.load(open(tmp_path / "x.pkl", "rb")).shape == (0, 4)
