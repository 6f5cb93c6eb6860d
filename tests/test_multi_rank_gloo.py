"""world_size-2 gloo test of the N>1 host logic (sharding by PRN / recording / channel + the result gather).
CPU only: the per-rank compute is replaced by deterministic fake results; the GPU box runs the same code over NCCL."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_result(prn):
    return {"found": 1, "doppler_bin": prn % 29, "code_phase_samples": 100 * prn, "carrier_freq": 500.0 * prn,
            "mag_relative": 1e6 + prn, "metric": 7.5 + prn}


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gnss_sdr_rs_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    present = {2, 3, 9, 14, 19, 32}
    base = 0xFFFFFFFF & ~(1 << 4)  # PRN 5 is already tracked: excluded from the search mask
    mask = sharding.prn_mask_for_rank(rank, world, 32, base)
    results = [(_fake_result(p + 1) if ((mask >> p) & 1 and (p + 1) in present) else None) for p in range(32)]
    gathered = sharding.all_gather_results(sharding.pack_results(results), dist)
    merged = sharding.merge_prn_shards(gathered)
    again = sharding.ResultGatherer(dist, None, 32).gather(results)   # the preallocated form bench.py uses
    assert all((a == b).all() for a, b in zip(gathered, again))
    # raw-struct form (what bench.py ships): the C-ABI result array goes over the wire as bytes
    import gnss_sdr_rs_b200._ffi as ffi
    rg = sharding.RawResultGatherer(dist, None, 32, ffi.AcqResult)
    for p in range(32):
        r = results[p]
        rg.results[p].prn = p + 1
        rg.results[p].found = 1 if r else 0
        rg.results[p].code_phase_samples = r["code_phase_samples"] if r else 0
        rg.results[p].carrier_freq = r["carrier_freq"] if r else 0.0
    per_rank = rg.gather()
    raw_found = sorted(int(x.prn) for arr in per_rank for x in arr if x.found)
    assert raw_found == sorted(present - {5}), raw_found
    for arr in per_rank:
        for x in arr:
            if x.found:
                assert x.code_phase_samples == 100 * x.prn and x.carrier_freq == 500.0 * x.prn
    np.save(os.path.join(out_dir, "merged_%d.npy" % rank), merged)
    np.save(os.path.join(out_dir, "mask_%d.npy" % rank), np.array([mask], np.uint64))
    # batch snapshot acquisition (BASELINE configs[4]): 13 recordings dealt to the ranks, one padded gather at the end
    mine = sharding.items_for_rank(13, rank, world)
    local = np.stack([sharding.pack_results([_fake_result(p + 1) if (p + i) % 5 == 0 else None for p in range(32)])
                      for i in mine])
    table = sharding.gather_batch(local, 13, dist)
    np.save(os.path.join(out_dir, "batch_%d.npy" % rank), table)
    # batch / channel partitioning
    items = list(sharding.items_for_rank(13, rank, world))
    np.save(os.path.join(out_dir, "items_%d.npy" % rank), np.array(items))
    dist.barrier()
    dist.destroy_process_group()


def test_prn_sharding_and_gather_world2(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    m0, m1 = np.load(tmp_path / "merged_0.npy"), np.load(tmp_path / "merged_1.npy")
    assert (m0 == m1).all()                                   # every rank ends with the same table
    found = {p + 1 for p in range(32) if m0[p, 0] == 1}
    assert found == {2, 3, 9, 14, 19, 32}
    for p in found:
        r = _fake_result(p)
        assert m0[p - 1].tolist() == [1.0, r["doppler_bin"], r["code_phase_samples"], r["carrier_freq"],
                                       r["mag_relative"], r["metric"]]
    k0, k1 = int(np.load(tmp_path / "mask_0.npy")[0]), int(np.load(tmp_path / "mask_1.npy")[0])
    assert k0 & k1 == 0 and (k0 | k1) == (0xFFFFFFFF & ~(1 << 4))   # disjoint cover of the search mask
    assert abs(bin(k0).count("1") - bin(k1).count("1")) <= 1
    i0, i1 = np.load(tmp_path / "items_0.npy"), np.load(tmp_path / "items_1.npy")
    assert sorted(i0.tolist() + i1.tolist()) == list(range(13))
    b0, b1 = np.load(tmp_path / "batch_0.npy"), np.load(tmp_path / "batch_1.npy")
    assert b0.shape == (13, 32, 6) and (b0 == b1).all()
    sys.path.insert(0, ROOT)
    from gnss_sdr_rs_b200 import sharding
    for i in range(13):     # recording order is preserved across the rank boundary
        want = sharding.pack_results([_fake_result(p + 1) if (p + i) % 5 == 0 else None for p in range(32)])
        assert (b0[i] == want).all()


def test_partition_edge_cases():
    sys.path.insert(0, ROOT)
    from gnss_sdr_rs_b200 import sharding
    for world in (1, 2, 4, 8):
        masks = [sharding.prn_mask_for_rank(r, world) for r in range(world)]
        acc = 0
        for m in masks:
            assert acc & m == 0
            acc |= m
        assert acc == 0xFFFFFFFF
        for n in (0, 1, 7, 512, 1024):
            items = [i for r in range(world) for i in sharding.items_for_rank(n, r, world)]
            assert items == list(range(n))
