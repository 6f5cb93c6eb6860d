"""Multi-GPU sharding of the hot path: the units are independent, so ranks never exchange
samples or spectra; the only collective is one small gather of per-PRN results (NCCL over NVLink on
the GPU box, gloo in the CPU tests).

  * acquisition of ONE recording: PRNs are dealt round-robin to ranks (`prn_mask_for_rank`), each rank
    runs the full Doppler grid for its PRNs, `all_gather_results` merges the per-PRN results;
  * batch acquisition: recordings are dealt to ranks (`items_for_rank`), results gathered at the end;
  * tracking: channels are dealt to ranks (`items_for_rank`), no collective during the run.
"""
import numpy as np

RESULT_FIELDS = ("found", "doppler_bin", "code_phase_samples", "carrier_freq", "mag_relative", "metric")


def prn_mask_for_rank(rank, world, n_prn=32, base_mask=0xFFFFFFFF):
    """Bit (prn-1) set for the PRNs this rank searches (the reference's mask convention, do_acquisition.rs:307)."""
    mask = 0
    k = 0
    for p in range(n_prn):
        if (base_mask >> p) & 1:
            if k % world == rank:
                mask |= 1 << p
            k += 1
    return mask


def items_for_rank(n_items, rank, world):
    """Contiguous block partition of recordings / channels."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return range(lo, min(n_items, lo + per))


def pack_results(results):
    """List (one per PRN) of result dicts / None -> float64 [n_prn, 6] tensor-ready array."""
    out = np.zeros((len(results), len(RESULT_FIELDS)), np.float64)
    for i, r in enumerate(results):
        if r:
            out[i] = [1.0, r["doppler_bin"], r["code_phase_samples"], r["carrier_freq"], r["mag_relative"],
                      r.get("metric", 0.0)]
    return out


def merge_prn_shards(packed_per_rank):
    """Every PRN is owned by exactly one rank: the merged table takes each row from the rank that found / searched it."""
    stack = np.stack(packed_per_rank)                      # [world, n_prn, 6]
    owner = stack[:, :, 0].argmax(axis=0)                  # rank that reports found=1 (or 0 if none)
    return stack[owner, np.arange(stack.shape[1])]


def all_gather_results(packed, dist, device=None):
    """One all_gather of a [n_prn, 6] array; returns the list of every rank's array."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(packed))
    if device is not None:
        t = t.to(device)
    bufs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(bufs, t)
    return [b.cpu().numpy() for b in bufs]


class ResultGatherer:
    """The per-search collective with preallocated buffers: packed results -> pinned host tensor -> device ->
    ONE all_gather_into_tensor over NCCL (gloo in CPU tests) -> host.  ~1.5 KB per rank: latency-bound."""

    def __init__(self, dist, device, n_prn=32):
        import torch
        self.dist, self.device, self.n_prn = dist, device, n_prn
        self.world = dist.get_world_size()
        cuda = device is not None and str(device).startswith("cuda")
        self.host_in = torch.zeros(n_prn, len(RESULT_FIELDS), dtype=torch.float64, pin_memory=cuda)
        self.dev_in = torch.zeros_like(self.host_in, device=device) if device is not None else self.host_in
        self.dev_out = torch.zeros(self.world * n_prn, len(RESULT_FIELDS), dtype=torch.float64,
                                   device=device if device is not None else "cpu")

    def gather(self, results):
        self.host_in.copy_(__import__("torch").from_numpy(pack_results(results)))
        if self.dev_in is not self.host_in:
            self.dev_in.copy_(self.host_in, non_blocking=True)
        self.dist.all_gather_into_tensor(self.dev_out, self.dev_in)
        out = self.dev_out.cpu().numpy().reshape(self.world, self.n_prn, len(RESULT_FIELDS))
        return [out[r] for r in range(self.world)]


def gather_batch(local_packed, n_total, dist, device=None):
    """Batch snapshot acquisition (BASELINE configs[4]): recordings are dealt to ranks with `items_for_rank`; each rank
    holds a [n_local, n_prn, 6] array of packed results.  ONE padded all_gather at the end returns the [n_total, n_prn, 6]
    table in recording order on every rank (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    per = (n_total + world - 1) // world
    local_packed = np.asarray(local_packed, np.float64)
    n_prn = local_packed.shape[1] if local_packed.ndim == 3 else 32
    pad = np.zeros((per, n_prn, len(RESULT_FIELDS)), np.float64)
    pad[:local_packed.shape[0]] = local_packed
    if dist is None or world == 1:
        return pad[:n_total]
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1, *t.shape[1:]), t)
    return out.cpu().numpy().reshape(world * per, n_prn, len(RESULT_FIELDS))[:n_total]


def search_batch(engine, recordings, num_integrations, rank=0, world=1, dist=None, device=None):
    """Runs `engine.search` on this rank's share of `recordings` (list of complex64 arrays, or a callable index -> array
    so that a rank only materialises its own) and gathers the per-recording, per-PRN results."""
    n_total = len(recordings) if not callable(recordings) else recordings(None)
    mine = items_for_rank(n_total, rank, world)
    local = [pack_results(engine.search(recordings(i) if callable(recordings) else recordings[i], num_integrations))
             for i in mine]
    local = np.stack(local) if local else np.zeros((0, engine.n_prn, len(RESULT_FIELDS)))
    return gather_batch(local, n_total, dist, device)


class RawResultGatherer:
    """The per-search collective on the C-ABI's own result structs: the (gb_acq_result * n_prn) array of every rank
    (1.8 KB) is viewed as bytes in a pinned host tensor, copied to the device and exchanged with ONE
    all_gather_into_tensor (NCCL over NVLink; gloo in the CPU tests) -- no per-field packing on the host."""

    def __init__(self, dist, device, n_prn, result_type):
        import ctypes
        import torch
        self.dist, self.device, self.n_prn, self.result_type = dist, device, n_prn, result_type
        self.world = dist.get_world_size()
        self.nbytes = ctypes.sizeof(result_type) * n_prn
        cuda = device is not None and str(device).startswith("cuda")
        self.host_in = torch.zeros(self.nbytes, dtype=torch.uint8, pin_memory=cuda)
        self.results = (result_type * n_prn).from_address(self.host_in.data_ptr())   # the search writes straight into it
        self.dev_in = torch.zeros_like(self.host_in, device=device) if cuda else self.host_in
        self.dev_out = torch.zeros(self.world * self.nbytes, dtype=torch.uint8, device=device if cuda else "cpu")
        self.host_out = torch.zeros(self.world * self.nbytes, dtype=torch.uint8, pin_memory=cuda)

    def gather(self):
        """Exchange self.results; returns a list (one per rank) of (result_type * n_prn) arrays viewing the host copy."""
        if self.dev_in is not self.host_in:
            self.dev_in.copy_(self.host_in, non_blocking=True)
        self.dist.all_gather_into_tensor(self.dev_out, self.dev_in)
        self.host_out.copy_(self.dev_out)
        base = self.host_out.data_ptr()
        return [(self.result_type * self.n_prn).from_address(base + r * self.nbytes) for r in range(self.world)]
