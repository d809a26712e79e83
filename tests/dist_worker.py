"""Worker of tests/test_dist_gpu.py: one rank per GPU (torchrun).  Every rank owns a contiguous shard of the streams
(bpv.dist.shard_range), runs the REAL engine on it and all-gathers the packed 24-byte records over NCCL; rank 0 also runs
the whole stream set on its own GPU and requires the gathered records to equal the single-GPU ones bit for bit
(SURVEY.md 4 iv / 8e).  Usage: torchrun ... tests/dist_worker.py <out_file> <config>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bp-from-video_b200'))
sys.path.insert(0, ROOT)


def main():
    from bpv import _cabi, dist as bdist, synth
    from bpv.engine import BatchedSignalProcessor
    out_file, config = sys.argv[1], sys.argv[2]
    rank, local, world = bdist.init_from_env('nccl')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    S = 37 if config == "c2" else 36                        # 37 streams: ragged shards (padded gather); 36: equal shards (overlapped gather)
    T, steps, W, H, Wd = 4, 14, 40, 48, 64
    cfg = dict(c2=dict(color_channel=_cabi.CHROM_GREEN, processing_methods=[_cabi.DETREND_LINEAR, _cabi.FILTER_FIR],
                       spectrum_transform=_cabi.PGRAM_WELCH),
               c5=dict(color_channel=_cabi.GREEN, processing_methods=[_cabi.FILTER_BUTTER], spectrum_transform=_cabi.PGRAM_LS))[config]
    rng = np.random.default_rng(2024)                        # the same data on every rank; each takes its shard
    frames = rng.integers(0, 256, (S, T * steps, H, Wd, 3), dtype=np.uint8)
    boxes = np.stack([synth.roi_boxes(rng, T * steps, H, Wd) for _ in range(S)])
    ts = np.stack([synth.timestamps(rng, T * steps, 30.0, irregular=True, drop=0.02, origin=rng.uniform(0, 5)) for _ in range(S)])

    def run(lo, hi):
        eng = BatchedSignalProcessor(hi - lo, 2, signal_max_samples=W, max_frames_per_step=T, device=dev, **cfg)
        recs = []
        for k in range(steps):
            sl = slice(k * T, (k + 1) * T)
            res = eng.step(torch.from_numpy(frames[lo:hi, sl]).to(dev), torch.from_numpy(boxes[lo:hi, sl]).to(dev),
                           torch.from_numpy(ts[lo:hi, sl].copy()).to(dev))
            recs.append(res.packed32().clone())
        return recs

    lo, hi = bdist.shard_range(S, rank, world)
    mine = run(lo, hi)
    counts = [(bdist.shard_range(S, r, world)[1] - bdist.shard_range(S, r, world)[0]) * T for r in range(world)]
    gathered = [bdist.gather_records(r, counts) for r in mine]
    # the overlapped gather (RecordGather) must deliver the same tensors one step later (equal shards only)
    ok_async = True
    if len(set(counts)) == 1:
        rg = bdist.RecordGather()
        for k, r in enumerate(mine):
            rg.launch(r)
            prev = rg.collect()
            if k >= 1:
                ok_async &= torch.equal(prev, gathered[k - 1])
        ok_async &= torch.equal(rg.flush(), gathered[-1])
    torch.cuda.synchronize()
    if rank == 0:
        single = run(0, S)
        ok = all(torch.equal(g, s) for g, s in zip(gathered, single))
        finite = float(torch.isfinite(single[-1].view(torch.float32)[:, :2]).float().mean())
        with open(out_file, 'w') as f:
            f.write(f'{"OK" if ok and ok_async else "MISMATCH"} world={world} steps={steps} records={single[-1].shape} finite_bpm={finite:.2f}\n')
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
