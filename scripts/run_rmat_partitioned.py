#!/usr/bin/env python
"""BASELINE configs[4] end to end on G GPUs of one box (torchrun): R-MAT scale S generated on
the devices, vertex-partitioned CSR with NVLink peer reads, num_walks x walk_length walks from
every vertex, then data-parallel SGNS with NCCL model averaging.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      scripts/run_rmat_partitioned.py --scale 26 --edge-factor 16 --num-walks 5 --walk-length 40 --dim 128

Prints one JSON line on rank 0 (also written to gpurun_out/ when that directory exists).
Every rank generates 1/G of the edges, both directions are routed to the owner of their source
vertex with one all-to-all, duplicates and self loops are dropped locally.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rmat_keys(scale, n_edges, seed, device, abcd=(0.57, 0.19, 0.19, 0.05), chunk=1 << 26):
    """n_edges R-MAT edges as int64 keys (src << 32 | dst), ids scrambled by an affine bijection."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a, b, c, _ = abcd
    n = 1 << scale
    mult = 0x9E3779B1 | 1                               # odd => bijection mod 2^scale
    out = []
    for lo in range(0, n_edges, chunk):
        m = min(chunk, n_edges - lo)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, device=device, generator=gen)
            src = (src << 1) | (r >= a + b).long()
            dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()
        src = (src * mult + 12345) & (n - 1)
        dst = (dst * mult + 12345) & (n - 1)
        out.append((src << 32) | dst)
        del src, dst, r
    return torch.cat(out)


def route_to_owners(keys, part_size, world):
    """All-to-all: every arc key goes to the rank owning its source vertex."""
    owner = torch.div(keys >> 32, part_size, rounding_mode="floor")
    order = torch.argsort(owner)
    keys = keys[order]
    send_counts = torch.bincount(owner, minlength=world)
    del owner, order
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    recv = torch.empty(int(recv_counts.sum()), dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(recv, keys, output_split_sizes=recv_counts.tolist(), input_split_sizes=send_counts.tolist())
    return recv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--num-walks", type=int, default=5)
    ap.add_argument("--walk-length", type=int, default=40)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--p", type=float, default=0.25)
    ap.add_argument("--q", type=float, default=4.0)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--no-sgns", action="store_true")
    ap.add_argument("--replicated-check", action="store_true", help="also walk a replicated copy and compare (small scales)")
    args = ap.parse_args()

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from node2vec_b200.graph import DeviceGraph, PartitionedGraph
    from node2vec_b200.sgns import Word2Vec

    V = 1 << args.scale
    S = (V + world - 1) // world
    n_edges = args.edge_factor << args.scale
    t_all = time.perf_counter()
    t0 = time.perf_counter()
    mine = n_edges * (rank + 1) // world - n_edges * rank // world
    keys = rmat_keys(args.scale, mine, 1000 + rank, dev)
    keys = keys[(keys >> 32) != (keys & 0xFFFFFFFF)]                        # no self loops
    keys = torch.cat([keys, ((keys & 0xFFFFFFFF) << 32) | (keys >> 32)])    # both directions
    keys = route_to_owners(keys, S, world)
    keys = torch.unique(keys)                                               # simple graph; sorted by (src, dst)
    src = (keys >> 32).to(torch.int32)
    dst = (keys & 0xFFFFFFFF).to(torch.int32)
    del keys
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0

    t0 = time.perf_counter()
    g = PartitionedGraph.from_local_arcs(src, dst, None, V, assume_symmetric=True)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    n_local_arcs = int(src.numel())
    if not args.replicated_check:
        del src, dst
    start = g.start_vertices()
    torch.cuda.empty_cache()

    # warm-up on a slice, then the timed walk over every start vertex of this rank
    g.walk(start[:1024], args.num_walks, args.walk_length, args.p, args.q, seed=1)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    walks, alive, _ = g.walk(start, args.num_walks, args.walk_length, args.p, args.q, seed=42)
    e1.record()
    torch.cuda.synchronize()
    walk_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    steps = torch.tensor([float(alive.sum().item()) * args.walk_length], device=dev, dtype=torch.float64)
    dist.all_reduce(walk_ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(steps)
    _, _, stats = g.walk(start[: min(int(start.numel()), 200000)], args.num_walks, args.walk_length, args.p, args.q,
                         seed=42, collect_stats=True)
    remote = (torch.div(walks[:100000, 1:], S, rounding_mode="floor") != rank).float().mean().item()

    check = None
    if args.replicated_check:
        all_src = [torch.empty(0, dtype=torch.int32, device=dev) for _ in range(world)]
        sizes = [None] * world
        dist.all_gather_object(sizes, n_local_arcs)
        all_src = [torch.empty(n, dtype=torch.int32, device=dev) for n in sizes]
        all_dst = [torch.empty(n, dtype=torch.int32, device=dev) for n in sizes]
        dist.all_gather(all_src, src)
        dist.all_gather(all_dst, dst)
        full = DeviceGraph.from_arcs(torch.cat(all_src), torch.cat(all_dst), None, n_vertices=V)
        ref, ref_alive, _ = full.walk(start, args.num_walks, args.walk_length, args.p, args.q, seed=42)
        check = bool(torch.equal(ref, walks) and torch.equal(ref_alive, alive))
        assert check, "partitioned walks differ from the replicated graph's"
        del full, ref

    out = {
        "what": "RMAT vertex-partitioned CSR + NVLink peer reads + DP SGNS (BASELINE configs[4] shape)",
        "n_gpus": world, "scale": args.scale, "vertices": V, "undirected_edges_generated": n_edges,
        "arcs": int(g.n_arcs), "arcs_per_gpu": n_local_arcs, "graph_bytes_per_gpu": int(g.nbytes()),
        "p": args.p, "q": args.q, "num_walks": args.num_walks, "walk_length": args.walk_length,
        "walk_steps": float(steps.item()), "walk_ms": float(walk_ms.item()),
        "walk_steps_per_s": float(steps.item()) / (float(walk_ms.item()) * 1e-3),
        "remote_hop_fraction": remote, "trials_per_step": stats["trials"] / max(stats["steps"], 1),
        "probes_per_step": stats["probes"] / max(stats["steps"], 1),
        "gen_s": t_gen, "build_s": t_build, "bit_identical_to_replicated": check,
    }
    if not args.no_sgns:
        walks = walks[alive] if not bool(alive.all()) else walks
        m = Word2Vec(size=args.dim, sg=1, negative=5, window=5, min_count=1, iter=args.epochs, seed=1,
                     batch_words=10000, process_group=dist.group.WORLD)
        t0 = time.perf_counter()
        m.build_vocab(walks)
        torch.cuda.synchronize()
        t_vocab = time.perf_counter() - t0
        dist.barrier()
        e0.record()
        m.train(walks)                      # epochs x (kernel + allreduce-average of both tables)
        e1.record()
        torch.cuda.synchronize()
        sg_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        pairs = torch.tensor([float(m.train_stats["pairs"])], device=dev, dtype=torch.float64)
        dist.all_reduce(sg_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pairs)
        out.update({"dim": args.dim, "sgns_epochs": args.epochs, "sgns_pairs": float(pairs.item()),
                    "sgns_ms": float(sg_ms.item()), "sgns_pairs_per_s": float(pairs.item()) / (float(sg_ms.item()) * 1e-3),
                    "sgns_table_bytes_per_gpu": int(2 * m.syn0.numel() * 4), "vocab_s": t_vocab,
                    "sgns_sync": "allreduce(sum)+scale of both tables every epoch (inside the timed region)"})
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    out["total_s"] = time.perf_counter() - t_all
    if rank == 0:
        line = json.dumps(out)
        print(line, flush=True)
        if os.path.isdir("gpurun_out"):
            with open(f"gpurun_out/rmat{args.scale}_{world}gpu.json", "w") as f:
                f.write(line + "\n")
    dist.barrier()
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
