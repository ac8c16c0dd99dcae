#!/usr/bin/env python
"""bench_step.py -- data-parallel training step (BASELINE configs 2 and 5): SDNetLite (whole DenseNet-121 siamese tower --
121 BatchNorm layers per pass --, 1x17 correlation + corrConv2d + ReLU, decoders, disparity warp + attention blend) fwd
+ CE/Lovasz/L1 loss + bwd + Adam on synthetic 256x512 stereo pairs, batch 4 per GPU, under DistributedDataParallel with
synchronised batch norm.  NCCL carries the gradient all-reduce; the BN statistics travel over NVLink peer memory inside
the bn_pair kernels (--nccl-bn: one NCCL collective per layer and direction instead).  One process per GPU; a single GPU
runs the same code path in a 1-rank process group, so the N = 1 point of a scaling curve is the same configuration:

    python bench_step.py --steps 30                                        # 1 GPU (config 2)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        bench_step.py --steps 30                                           # config 5 (global batch 32)

Device time of K steps between barriers, max over ranks; rank 0 prints one JSON line (pairs/s over all ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import torch

    from pmt_learning_for_semantic_segmentation_and_disparity_b200 import harness, sharding

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--no-sync-bn", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--no-pair", action="store_true",
                    help="run the siamese tower twice (one SyncBatchNorm collective per call) instead of once over "
                         "[left; right] with per-half statistics and one collective per layer")
    ap.add_argument("--json-out", default="", help="also write the JSON record to this file (rank 0)")
    ap.add_argument("--nccl-bn", action="store_true", help="exchange the BN statistics with NCCL collectives (one all_gather / "
                                                           "all_reduce launch per layer and direction) instead of the "
                                                           "NVLink peer-memory exchange inside the bn_pair kernels")
    ap.add_argument("--shallow", action="store_true", help="stop the tower after denseblock2 (39 BN layers) instead of running "
                                                           "the whole DenseNet-121 per image (121 BN layers per tower pass, the "
                                                           "reference's `densenet` backbone)")
    ap.add_argument("--unfused", action="store_true", help="sampler -> conv -> relu and warp -> blend as separate ops")
    ap.add_argument("--no-lovasz", action="store_true", help="drop the Lovasz-Softmax term of the seg2 loss")
    ap.add_argument("--plain-single", action="store_true", help="at 1 GPU run WITHOUT a process group (no DDP, stock BatchNorm); "
                                                                "default: a 1-rank group, i.e. the N-GPU code path")
    ap.add_argument("--graph", action="store_true", help="(default) capture the whole step, NCCL collectives included, "
                                                         "in one CUDA graph")
    args = ap.parse_args()
    force_group = not args.plain_single
    if not args.eager and (int(os.environ.get("WORLD_SIZE", "1")) > 1 or force_group):
        # capturing NCCL collectives: the watchdog must not query/abort work that lives inside a capture
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        os.environ.setdefault("TORCH_NCCL_ENABLE_MONITORING", "0")
        if os.environ.get("PMT_STEP_HANG_DUMP"):
            import faulthandler

            faulthandler.dump_traceback_later(int(os.environ["PMT_STEP_HANG_DUMP"]), exit=True)
    world = sharding.init_world("nccl" if (int(os.environ.get("WORLD_SIZE", "1")) > 1 or force_group) else None,
                                force_group=force_group)
    dev = torch.device("cuda", world.local_rank)
    torch.cuda.set_device(dev)
    use_graph = not args.eager
    paired = not args.no_pair and not args.no_sync_bn
    peer = paired and not args.nccl_bn and world.distributed
    step, model = harness.build_training_step(world, batch_per_gpu=args.batch, sync_bn=not args.no_sync_bn,
                                               cuda_graph=use_graph, paired_tower=not args.no_pair, peer_bn=peer,
                                               full_depth=not args.shallow, fused_ops=not args.unfused,
                                               lovasz=not args.no_lovasz)
    for _ in range(max(args.warmup, 3)):
        loss = step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sharding.barrier(world)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize(dev)
    sharding.barrier(world)
    value, ms = sharding.throughput(world, args.batch * args.steps, e0.elapsed_time(e1))
    if world.is_main:
        n_params = sum(p.numel() for p in model.parameters())
        n_bn_paired = sum(isinstance(m, harness.PairedSyncBatchNorm) for m in model.modules())
        n_bn_sync = sum(isinstance(m, torch.nn.SyncBatchNorm) for m in model.modules())
        rec = {"metric": "SDNetLite training step pairs/s (256x512, batch 4/GPU, DDP+SyncBN)", "value": value,
               "unit": "pairs/s", "n_gpus": world.world_size, "steps": args.steps,
               "ms_per_step": ms / args.steps, "scaling": "weak", "global_batch": args.batch * world.world_size,
               "params": n_params, "loss": float(loss.detach()), "sync_bn": not args.no_sync_bn,
               "cuda_graph": use_graph, "paired_tower": paired, "bn_exchange": "nvlink-peer" if peer else "nccl",
               "full_depth": not args.shallow, "fused_ops": not args.unfused, "lovasz": not args.no_lovasz,
               "bn_layers_paired": n_bn_paired, "bn_layers_sync": n_bn_sync,
               "nccl_launches_per_step_bn": (0 if peer else 2 * n_bn_paired) + 2 * n_bn_sync if world.distributed else 0,
               "collectives": "DDP gradient all-reduce (NCCL); BN statistics: " +
                              ("NVLink peer-memory exchange inside the bn_pair kernels" if peer else "NCCL all_gather / all_reduce")}
        print(json.dumps(rec), flush=True)
        if args.json_out:
            with open(args.json_out, "w") as f:
                json.dump(rec, f)
    if getattr(step, "exchange", None) is not None:
        step.exchange.check()      # a kernel that gave up waiting for a peer would have produced garbage
    if use_graph and world.distributed:
        # A CUDA graph that holds captured NCCL kernels keeps the communicator busy: destroy_process_group() was
        # observed to block forever behind it (that -- not the capture -- was the "hang" of the first attempts).
        # Release the graph first, and never let teardown outlive the measurement.
        import threading

        sys.stdout.flush()
        step.graph.reset()
        torch.cuda.synchronize(dev)
        t = threading.Thread(target=sharding.shutdown, args=(world,), daemon=True)
        t.start()
        t.join(15.0)
        if t.is_alive():
            os._exit(0)
        return
    sharding.shutdown(world)


if __name__ == "__main__":
    main()
