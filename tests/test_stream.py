"""Streaming driver of BASELINE configs[3] (stream.py): host logic on CPU with a stand-in handle (two ranks over gloo),
and on the GPU the device generator + the double-buffered loop against per-image runs and the oracle."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

from multimodal_biometric_fingerprints_palms_b200 import synth
from multimodal_biometric_fingerprints_palms_b200.stream import batch_plan, run_stream


def test_batch_plan_covers_slice_in_order():
    assert batch_plan(5, 5, 4) == []
    assert batch_plan(0, 10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert batch_plan(7, 9, 100) == [(7, 2)]
    with pytest.raises(ValueError):
        batch_plan(0, 3, 0)


def test_philox_known_answers():
    # Random123 known-answer vectors of Philox4x32-10
    assert [int(v) for v in synth.philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(v) for v in synth.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == \
        [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert [int(v) for v in synth.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)] == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_counter_image_is_a_ridge_print_and_index_addressable():
    a = synth.ridge_image_counter(320, 240, seed=5, index=123456789012)
    b = synth.ridge_image_counter(320, 240, seed=5, index=123456789012)
    c = synth.ridge_image_counter(320, 240, seed=5, index=123456789013)
    assert a.dtype == np.uint8 and a.shape == (320, 240) and np.array_equal(a, b) and not np.array_equal(a, c)
    assert a[:4].mean() > 200 and 100 < a[100:220, 60:180].mean() < 170       # bright background, ridges inside the ellipse


class _FakePipe:
    """Stands in for FingerprintPipeline without a GPU: 'refined count' of image i is i % 7."""
    def __init__(self, H, W, max_batch):
        self.max_batch, self.last_n, self.first, self.closed = max_batch, 0, None, False

    def synth_ridge(self, seed, first, n, period=0.0, noise_sigma=12.0):
        assert 1 <= n <= self.max_batch
        self.first, self.n = first, n

    def run_input_async(self, n):
        assert n == self.n
        self.last_n = n

    def download_refined(self):
        pass

    def result_block(self, cap=64):
        idx = np.arange(self.first, self.first + self.last_n)
        roi = np.stack([idx, idx, idx, idx], 1).astype(np.int32)
        return roi, (idx % 11).astype(np.int32), (idx % 7).astype(np.int32), np.zeros((self.last_n, cap), np.uint8)

    def close(self):
        self.closed = True


def test_run_stream_single_rank_order_and_counts():
    seen = []
    st = run_stream(103, 1, _FakePipe, batch=10, on_batch=lambda first, roi, rc, oc, ref: seen.append((first, len(oc), int(roi[0, 0]))))
    assert [s[0] for s in seen] == list(range(0, 103, 10)) and all(s[0] == s[2] for s in seen)
    assert st["images"] == 103 and st["batches"] == 11 and st["refined"] == sum(i % 7 for i in range(103))
    assert run_stream(0, 1, _FakePipe)["batches"] == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = run_stream(total, 9, _FakePipe, batch=16, rank=rank, world=world)
    t = torch.tensor([st["images"], st["refined"], st["raw"]], dtype=torch.int64)
    dist.all_reduce(t)                          # only counters cross ranks - no data-path collective
    if rank == 0:
        q.put(t.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_run_stream_world_size_2_gloo():
    total = 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [total, sum(i % 7 for i in range(total)), sum(i % 11 for i in range(total))]


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_device_generator_matches_numpy_twin_within_one_level():
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    p = FingerprintPipeline(320, 240, max_batch=6)
    p.synth_ridge(77, 4_000_000_000, 6)                   # indices beyond 2^32: the 64-bit counter path
    dev = p.fetch_input(6)
    worst = 0
    for i in range(6):
        twin = synth.ridge_image_counter(320, 240, seed=77, index=4_000_000_000 + i)
        d = np.abs(dev[i].astype(np.int16) - twin.astype(np.int16))
        worst = max(worst, int(d.max()))
        assert (d > 0).mean() < 0.02, f"image {i}: {(d > 0).mean():.4f} of the pixels differ"
    assert worst <= 1, worst
    q = FingerprintPipeline(200, 184, max_batch=2)        # odd width path (W/2 pairs per row)
    q.synth_ridge(3, 10, 2, period=9.0, noise_sigma=0.0)
    tw = synth.ridge_image_counter(200, 184, seed=3, index=11, period=9.0, noise_sigma=0.0)
    assert np.abs(q.fetch_input(2)[1].astype(np.int16) - tw.astype(np.int16)).max() <= 1


@pytest.mark.gpu
def test_stream_equals_per_image_runs_and_oracle():
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    from oracle import ref_pipeline as rp
    total, seed = 150, 2024
    got = {}

    def keep(first, roi, rc, oc, ref):
        for i in range(len(oc)):
            got[first + i] = (tuple(int(v) for v in roi[i]), int(rc[i]), [(int(m["x"]), int(m["y"]), int(m["type"]), float(m["quality"]))
                                                                          for m in ref[i, :oc[i]]])

    mk = lambda H, W, B: FingerprintPipeline(H, W, max_batch=B)
    stats = [run_stream(total, seed, mk, batch=32, rank=r, world=2, on_batch=keep) for r in range(2)]   # two "ranks" in turn
    assert sorted(got) == list(range(total)) and sum(s["images"] for s in stats) == total
    assert sum(s["refined"] for s in stats) == sum(len(v[2]) for v in got.values())
    one = FingerprintPipeline(320, 240, max_batch=1)
    for idx in (0, 31, 32, 74, 75, 149):                 # batch / rank boundaries: any index can be regenerated alone
        one.synth_ridge(seed, idx, 1)
        img = one.fetch_input(1)[0]
        one.run_decoded(1)
        mine = [(m["x"], m["y"], 0 if m["type"] == "ending" else 1, m["quality"]) for m in one.minutiae(0)]
        assert got[idx][0] == one.roi(0) and got[idx][2] == mine, idx
        ref = rp.enhance_to_minutiae(img, handoff="file")
        want = [(m["x"], m["y"], 0 if m["type"] == "ending" else 1) for m in ref["minutiae"]]
        assert [m[:3] for m in mine] == want, idx
        assert got[idx][1] == len(ref["raw_minutiae"])
