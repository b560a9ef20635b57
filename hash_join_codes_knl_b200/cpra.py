"""CPRA across GPUs: the reference's chunk-local partitioning + per-owner gather
(cpra2.cpp:1783-1827 local passes, :1868-1906 / :1940-1959 the `memcpy` gather timed as
"copy:") with threads replaced by GPUs -- one process per GPU, torch.distributed for the
plumbing.

  1. split     GPU g radix-partitions ITS chunk of R and S by owner = top log2(G) bits of
               key*factor (Engine.cpra_split, kernels of csrc/radix.cu)
  2. exchange  per-owner counts (all_to_all of G int64), then one variable-size all-to-all per
               column over NVLink (NCCL); this replaces the reference's remote memcpy gather
  3. join      every GPU joins what it received with the PHJ kernels below the owner bits
               (Engine.cpra_join_local)
  4. reduce    count and the three checksums are summed over ranks (all_reduce)

`split_fn` / `join_fn` are injectable so the exchange logic can be exercised on CPU tensors
with the gloo backend (tests/test_cpra_gloo.py); the product path always passes an Engine."""
import torch
import torch.distributed as dist


def exchange_columns(columns, offsets, group=None):
    """All-to-all of tuple columns grouped by owner.

    columns: list of 1-D tensors of equal length, rows [offsets[g], offsets[g+1]) go to rank g.
    Returns (received columns, recv_counts)."""
    world = dist.get_world_size(group)
    dev = columns[0].device
    send_counts = torch.tensor([offsets[g + 1] - offsets[g] for g in range(world)], dtype=torch.int64, device=dev)
    recv_counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send_list = [int(x) for x in send_counts.tolist()]
    recv_list = [int(x) for x in recv_counts.tolist()]
    total = sum(recv_list)
    out = []
    for col in columns:
        recv = torch.empty(total, dtype=col.dtype, device=dev)
        dist.all_to_all_single(recv, col, output_split_sizes=recv_list, input_split_sizes=send_list, group=group)
        out.append(recv)
    return out, recv_list


def reduce_checks(count, sum_key, sum_outer, sum_inner, device, group=None):
    """Sum (count, 3 checksums) over ranks with uint64 wrap-around (int64 adds have the same bits)."""
    def to_i64(x):
        x &= (1 << 64) - 1
        return x - (1 << 64) if x >= (1 << 63) else x
    t = torch.tensor([to_i64(count), to_i64(sum_key), to_i64(sum_outer), to_i64(sum_inner)], dtype=torch.int64,
                     device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(int(x) & ((1 << 64) - 1) for x in t.tolist())


def cpra_join(engine, inner_chunk, outer_chunk, group=None, split_fn=None, join_fn=None, **opts):
    """Joins the union of all ranks' chunks.  Returns a dict with the GLOBAL count / checksums,
    this rank's JoinResult (`local`: its share of the rows stays on its GPU, like the per-thread
    output blocks of the reference) and per-step device times in ms."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    split_fn = split_fn or (lambda r, s, g: engine.cpra_split(r, s, g, **opts))
    join_fn = join_fn or (lambda r, s, me, g: engine.cpra_join_local(r, s, me, g, **opts))
    sp = split_fn(inner_chunk, outer_chunk, world)
    dev = sp["r_keys"].device
    use_events = dev.type == "cuda"
    if use_events:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
    if world == 1:
        rk, rv, sk, sv = sp["r_keys"], sp["r_vals"], sp["s_keys"], sp["s_vals"]
    else:
        (rk, rv), _ = exchange_columns([sp["r_keys"], sp["r_vals"]], sp["r_offsets"], group)
        (sk, sv), _ = exchange_columns([sp["s_keys"], sp["s_vals"]], sp["s_offsets"], group)
    exchange_ms = 0.0
    if use_events:
        ev[1].record()
        torch.cuda.current_stream(dev).synchronize()
        exchange_ms = ev[0].elapsed_time(ev[1])
    local = join_fn((rk, rv), (sk, sv), rank, world)
    count, sum_key, sum_outer, sum_inner = reduce_checks(local.count, local.sum_key, local.sum_outer,
                                                         local.sum_inner, dev, group)
    return {"count": count, "sum_key": sum_key, "sum_outer": sum_outer, "sum_inner": sum_inner, "local": local,
            "split_ms": float(sp.get("ms", 0.0)), "exchange_ms": exchange_ms,
            "join_ms": float(getattr(local, "seconds", 0.0)) * 1e3,
            "recv_tuples": (int(rk.numel()), int(sk.numel()))}


class FusedExchange:
    """State of the fused GPU-assign + exchange pass for one Engine / process group: this GPU's
    receive buffers, the peers' buffers mapped through CUDA IPC, and the small device tensors
    the collectives work on.  (Re)built collectively whenever some rank would receive more rows
    than the current capacity."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.r_cap = self.s_cap = 0
        self.own = None
        self.peers = None          # [column][owner] device pointers valid in this process
        self._opened = []
        dev = torch.device(f"cuda:{engine.device}")
        self.counts = torch.zeros(2 * self.world, dtype=torch.int64, device=dev)
        self.matrix = torch.zeros(2 * self.world * self.world, dtype=torch.int64, device=dev)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)
        self.token2 = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sums = torch.zeros(4, dtype=torch.int64, device=dev)
        # staged exchange: the plan of a step (made once for this state), its count tensors, the copies' side stream
        self.stage_plan = None
        self.stage_counts = self.stage_matrix = None
        self.side = torch.cuda.Stream(device=dev, priority=-1)

    def ensure(self, r_need, s_need):
        """r_need / s_need: the largest row count any rank receives (identical on every rank)."""
        if r_need <= self.r_cap and s_need <= self.s_cap:
            return
        for p in self._opened:
            self.engine.ipc_close(p)
        self._opened = []
        if self.own is not None:
            dist.barrier(group=self.group)       # nobody still maps the buffers about to be freed
        self.r_cap = max(self.r_cap, int(r_need * 1.05) + 1024)
        self.s_cap = max(self.s_cap, int(s_need * 1.05) + 1024)
        self.own = self.engine.cpra_recv_alloc(self.r_cap, self.s_cap)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.own["ipc"], group=self.group)
        self.peers = [[None] * self.world for _ in range(4)]
        for g in range(self.world):
            for c in range(4):
                if g == self.rank:
                    self.peers[c][g] = self.own["ptrs"][c]
                else:
                    self.peers[c][g] = self.engine.ipc_open(handles[g][c])
                    self._opened.append(self.peers[c][g])
        self.engine.cpra_bind(self.rank, self.world, self.peers, self.r_cap, self.s_cap)


MAX_HOT_KEYS = 256          # hjb200.h: hjb_cpra_split_hot
MAX_HOT_BUILD = 4096        # build tuples with hot keys, all ranks together (hjb_cpra_hot_join)


def detect_hot_keys(outer_keys, group=None, sample=1 << 16, per_rank=64, min_share=1.0 / 2048):
    """Heavy hitters of the probe side, the same list on every rank: every rank counts the keys of a strided
    sample of its chunk, nominates its most frequent ones, the nominations are all-gathered and a key is hot
    when its sampled frequency over all ranks reaches `min_share` of the sampled tuples.  Returns a sorted
    int32 CUDA tensor of at most MAX_HOT_KEYS keys (empty: no skew worth handling)."""
    world = dist.get_world_size(group)
    dev = outer_keys.device
    n = int(outer_keys.numel())
    m = min(sample, n)
    nominations = torch.zeros(2, per_rank, dtype=torch.int64, device=dev)
    if m:
        picks = outer_keys[:: max(1, n // m)][:m]
        keys, counts = torch.unique(picks, return_counts=True)
        top = torch.topk(counts, min(per_rank, int(counts.numel())))
        nominations[0, : top.indices.numel()] = keys[top.indices].to(torch.int64) & 0xFFFFFFFF
        nominations[1, : top.indices.numel()] = top.values
    gathered = torch.empty(world, 2, per_rank, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1), nominations.view(-1), group=group)
    keys, inverse = torch.unique(gathered[:, 0, :].reshape(-1), return_inverse=True)
    totals = torch.zeros(keys.numel(), dtype=torch.int64, device=dev).index_add_(0, inverse, gathered[:, 1, :].reshape(-1))
    sampled = torch.tensor([m], dtype=torch.int64, device=dev)
    dist.all_reduce(sampled, group=group)
    hot = (totals >= max(8, int(int(sampled.item()) * min_share))) & (keys != 0xFFFFFFFF)     # 0xFFFFFFFF is never declared hot
    keys, totals = keys[hot], totals[hot]
    if keys.numel() > MAX_HOT_KEYS:
        keys = keys[torch.topk(totals, MAX_HOT_KEYS).indices]
    keys = torch.sort(keys).values
    return (keys - ((keys >= (1 << 31)).to(torch.int64) << 32)).to(torch.int32).contiguous()     # uint32 bit pattern in int32


def split_hot(engine, inner_chunk, outer_chunk, hot_keys, group=None):
    """-> (cold outer chunk, this rank's hot outer tuples, all ranks' hot inner tuples) or None when the hot build
    tuples do not fit (then the step runs without the hot path)."""
    world = dist.get_world_size(group)
    dev = hot_keys.device
    mine_k = torch.zeros(MAX_HOT_BUILD, dtype=torch.int32, device=dev)
    mine_v = torch.zeros(MAX_HOT_BUILD, dtype=torch.int32, device=dev)
    found = engine.cpra_select_hot(inner_chunk, hot_keys, mine_k, mine_v)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([found], dtype=torch.int64, device=dev), group=group)
    counts = [int(x) for x in counts.tolist()]
    if sum(counts) > MAX_HOT_BUILD:
        return None
    width = max(1, max(counts))
    allk = torch.empty(world * width, dtype=torch.int32, device=dev)
    allv = torch.empty(world * width, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(allk, mine_k[:width].contiguous(), group=group)
    dist.all_gather_into_tensor(allv, mine_v[:width].contiguous(), group=group)
    keep = torch.cat([torch.arange(width, device=dev) < c for c in counts])
    hot_inner = (allk[keep].contiguous(), allv[keep].contiguous())
    cold, hot_outer = engine.cpra_split_hot(outer_chunk, hot_keys)
    return cold, hot_outer, hot_inner


def cpra_join_fused(engine, inner_chunk, outer_chunk, state, group=None, skew=False, **opts):
    """CPRA with the all-to-all fused into the GPU-assign pass, the whole step enqueued on ONE stream:
    count -> all-gather of the count matrix (NCCL) -> bases on the device -> every sender scatters
    straight into the owners' receive buffers over NVLink -> 1-element all-reduce (every sender's stores
    have landed) -> local join with the received sizes read on the device -> all-reduce of the
    checksums.  The host synchronises once, at the end.  Same result dict as cpra_join.

    skew=True: the probe tuples of heavy-hitter keys (detect_hot_keys) stay on their rank and are joined there
    against the replicated build tuples of those keys (hjb_cpra_split_hot / _select_hot / _hot_join), the rest
    takes the normal path -- the owners then receive balanced shares (the reference's static
    ownership cpra2.cpp:1868-1872 does not).  Device columns only.

    Requires the Engine to run on torch's current stream (Engine(use_torch_stream=True)): NCCL's
    collectives are ordered with the library's kernels through that stream."""
    world, rank = state.world, state.rank
    from .api import HjbCapacityError
    size = lambda col: int(col.numel()) if hasattr(col, "numel") else int(col.size)
    hot_parts, n_hot = None, 0
    if skew:
        hot_keys = detect_hot_keys(outer_chunk[0], group)
        n_hot = int(hot_keys.numel())
        if n_hot:
            hot_parts = split_hot(engine, inner_chunk, outer_chunk, hot_keys, group)
            if hot_parts is not None:
                outer_chunk = hot_parts[0]
    nr, ns = size(inner_chunk[0]), size(outer_chunk[0])
    if state.own is None:
        # first step: room for a uniform share plus a quarter; a skewed input grows it below
        tot = torch.tensor([nr, ns], dtype=torch.int64, device=state.counts.device)
        dist.all_reduce(tot, group=group)
        tr, ts = (int(x) for x in tot.tolist())
        state.ensure(tr // world + tr // (4 * world) + 1024, ts // world + ts // (4 * world) + 1024)
    while True:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        engine.cpra_count_async(inner_chunk, outer_chunk, state.counts, **opts)
        dist.all_gather_into_tensor(state.matrix, state.counts, group=group)
        engine.cpra_scatter_async(state.matrix)
        ev[1].record()
        dist.all_reduce(state.token, group=group)            # completes here once every rank has passed its scatter
        ev[2].record()
        engine.cpra_join_async(r_expect=state.expect[0] if hasattr(state, "expect") else 0,
                               s_expect=state.expect[1] if hasattr(state, "expect") else 0, **opts)
        if hot_parts is not None:
            engine.cpra_hot_join(hot_parts[1], hot_parts[2])
        state.sums.copy_(engine.cpra_sums_dev())
        dist.all_reduce(state.sums, group=group)
        ev[3].record()
        try:
            local, received, largest = engine.cpra_finish()
        except HjbCapacityError as e:
            state.ensure(*e.largest)                         # every rank saw the same matrix and takes this branch
            continue
        break
    state.expect = received
    count, sum_key, sum_outer, sum_inner = (int(x) & ((1 << 64) - 1) for x in state.sums.tolist())
    return {"count": count, "sum_key": sum_key, "sum_outer": sum_outer, "sum_inner": sum_inner, "local": local,
            "split_ms": ev[0].elapsed_time(ev[1]), "exchange_ms": ev[1].elapsed_time(ev[2]),
            "join_ms": float(local.seconds) * 1e3, "step_ms": ev[0].elapsed_time(ev[3]),
            "recv_tuples": received, "largest_recv": largest, "hot_keys": n_hot if hot_parts is not None else 0,
            "hot_outer_tuples": int(hot_parts[1][0].numel()) if hot_parts is not None else 0}


def cpra_join_staged(engine, inner_chunk, outer_chunk, state, group=None, overlap=True, parts=None, **opts):
    """CPRA with the STAGED exchange (csrc/stage.cu; the reference's own order -- chunk-local passes, then the gather of
    whole partition pieces, cpra2.cpp:1783-1827,1861-1905): stage A partitions the chunk locally by owner and
    sub-partition; the runs then leave in `parts` pieces (ranges of sub-partitions, R's piece before S's) as copies on a
    high-priority side stream, and while piece k + 1 crosses NVLink the owners already run the local pass over piece
    k and join its partitions.  Two radix passes over the data where the fused path needs three.  The plan (how the
    radix bits are split, how many parts) is made once per `state` from the first step's sizes; inputs that need more
    than two 9-bit passes take cpra_join_fused.  Same result dict as cpra_join_fused.

    overlap=False runs the copies on the main stream (no side stream): for A/B runs."""
    world = state.world
    from .api import HjbCapacityError
    size = lambda col: int(col.numel()) if hasattr(col, "numel") else int(col.size)
    dev = state.counts.device
    if state.stage_plan is None:
        tot = torch.tensor([size(inner_chunk[0]), size(outer_chunk[0])], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=group)
        tr, ts = (int(x) for x in tot.tolist())
        plan = engine.cpra_stage_plan(world, max(1, tr // world), max(1, ts // world), **opts)
        state.stage_plan = plan if plan is not None else "fused"
        if plan is not None:
            f = 1 << plan[0]
            state.stage_counts = torch.zeros(2 * f, dtype=torch.int64, device=dev)
            state.stage_matrix = torch.zeros(2 * f * world, dtype=torch.int64, device=dev)
            # pieces worth pipelining: a few hundred microseconds of copying each
            import os
            auto = int(os.environ.get("HJB_STAGE_PARTS", "0")) or (4 if (tr + ts) // world >= (1 << 26) else 1)
            state.stage_parts = max(1, min(parts or auto, f // world, 8))
            state.stage_tokens = torch.zeros(2 * state.stage_parts, dtype=torch.int32, device=dev)
        if state.own is None:
            state.ensure(tr // world + tr // (4 * world) + 1024, ts // world + ts // (4 * world) + 1024)
    if state.stage_plan == "fused":
        return cpra_join_fused(engine, inner_chunk, outer_chunk, state, group, **opts)
    abits, bbits, big_fill = state.stage_plan
    K = state.stage_parts
    main = torch.cuda.current_stream(dev)
    side = state.side if overlap else main
    # the order the pieces leave in: R's piece one ahead of S's, so that S's stage A has finished when its first piece is due
    order = [(0, 0)]
    for k in range(K):
        if k + 1 < K:
            order.append((0, k + 1))
        order.append((1, k))
    while True:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        arrived = {piece: torch.cuda.Event(enable_timing=True) for piece in order}
        through = {}
        ev[0].record()
        engine.cpra_stage_count_async(inner_chunk, outer_chunk, abits, state.stage_counts, nparts=K, **opts)
        dist.all_gather_into_tensor(state.stage_matrix, state.stage_counts, group=group)
        engine.cpra_stage_scatter_async(state.stage_matrix, 0)
        ev[1].record()                                       # R is staged
        engine.cpra_stage_scatter_async(state.stage_matrix, 1)
        ev[2].record()                                       # S is staged
        with torch.cuda.stream(side):
            side.wait_event(ev[1])
            s_waited = False
            for i, (rel, k) in enumerate(order):
                if rel == 1 and not s_waited:
                    side.wait_event(ev[2])
                    s_waited = True
                engine.cpra_stage_copy_async(rel, side.cuda_stream, part=k)
                arrived[(rel, k)].record()                   # this rank's copies of the piece are through
                # completes once EVERY rank's copies of the piece are through; asynchronous: the side stream goes straight on to
                # the next piece, only the main stream will wait for it
                through[(rel, k)] = dist.all_reduce(state.stage_tokens[i:i + 1], group=group, async_op=True)
        for k in range(K):
            through[(0, k)].wait()                           # orders the main stream behind the collective, not the host
            engine.cpra_stage_local_async(bbits, big_fill, 0, part=k, **opts)
            through[(1, k)].wait()
            engine.cpra_stage_local_async(bbits, big_fill, 1, part=k, **opts)
        state.sums.copy_(engine.cpra_sums_dev())
        dist.all_reduce(state.sums, group=group)
        ev[3].record()
        try:
            local, received, largest = engine.cpra_finish()
        except HjbCapacityError as e:
            state.ensure(*e.largest)                         # every rank saw the same matrix and takes this branch
            continue
        break
    state.expect = received
    count, sum_key, sum_outer, sum_inner = (int(x) & ((1 << 64) - 1) for x in state.sums.tolist())
    last = arrived[order[-1]]
    return {"count": count, "sum_key": sum_key, "sum_outer": sum_outer, "sum_inner": sum_inner, "local": local,
            "split_ms": ev[0].elapsed_time(ev[2]), "exchange_ms": ev[1].elapsed_time(last),
            "join_ms": float(local.seconds) * 1e3, "step_ms": ev[0].elapsed_time(ev[3]),
            "copy_r_done_ms": ev[0].elapsed_time(arrived[(0, K - 1)]), "copy_s_done_ms": ev[0].elapsed_time(last),
            "recv_tuples": received, "largest_recv": largest, "hot_keys": 0, "hot_outer_tuples": 0,
            "stage_plan": {"stage_a_bits": abits, "local_bits": bbits, "big_fill": big_fill, "parts": K}}
