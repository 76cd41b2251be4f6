"""MNIST MLP on the fabric parameter-server engine (the B200 fast path), between-graph or in-graph.

between-graph (one process per GPU; rank 0..num_ps-1 are ps shards, the rest workers):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/fabric_mnist.py --num_ps 1 --issync

in-graph (ONE client process drives every GPU, like reference example_in_graph.py but with parameters sharded
over `--num_ps` ps GPUs and one training replica per remaining GPU):

    python examples/fabric_mnist.py --in_graph --gpus 8 --num_ps 2 --issync --optimizer adam

Same flags as distributed_mnist.py where they make sense; checkpoints are written by the chief (ps shard 0's
process) in the SAME name-keyed format, so examples/distributed_mnist_predict.py restores them.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.parallel.fabric import Fabric
from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
from distributed_tensorflow_b200.train.saver import update_checkpoint_state, write_bundle
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist

flags = dtf.app.flags
flags.DEFINE_integer("hidden_units", 100, "hidden layer width")
flags.DEFINE_integer("train_steps", 2000, "global steps to run (sync: aggregates; async: applies / num_workers)")
flags.DEFINE_integer("batch_size", 100, "per-worker batch")
flags.DEFINE_float("learning_rate", 0.01, "learning rate")
flags.DEFINE_string("optimizer", "adam", "sgd | momentum | adam")
flags.DEFINE_bool("issync", True, "synchronous replicas")
flags.DEFINE_integer("num_ps", 1, "ps shards")
flags.DEFINE_bool("in_graph", False, "one process drives all GPUs")
flags.DEFINE_integer("gpus", 0, "GPUs to use in in-graph mode (0: all)")
flags.DEFINE_bool("ps_on_workers", False, "every GPU runs a worker; ps shard s shares worker s's GPU and stream (no ps-only GPU)")
flags.DEFINE_string("train_dir", "/tmp/dtf_ckpt/fabric_mnist", "checkpoint directory")
flags.DEFINE_integer("num_train", 55000, "synthetic train-set size")
flags.DEFINE_string("input", "device", "device: the train split lives in every worker's HBM; host: every step's batch is copied from "
                                        "pinned host memory and every step's loss is read back (PSTrainEngine.train_loop: the reference's "
                                        "feed_dict loop, K steps per native call)")
flags.DEFINE_string("nvls", "auto", "NVLS multicast fabric: off | on | auto (auto = on when the box has NVLS and one process "
                                    "per GPU is used; the one-process topology opts in with 'on')")
FLAGS = flags.FLAGS


def main():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if FLAGS.in_graph or world == 1:
        n = FLAGS.gpus or torch.cuda.device_count()
        fabric = Fabric(n, {r: r for r in range(n)})
        rank = 0
    else:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
        fabric = Fabric.from_torch_distributed()
        n, rank = world, dist.get_rank()
    colocated = n == 1
    pow_ = FLAGS.ps_on_workers and not colocated
    cfg = EngineConfig(num_ps=1 if colocated else FLAGS.num_ps, num_workers=1 if colocated else (n if pow_ else n - FLAGS.num_ps),
                       sync=FLAGS.issync, colocated=colocated, ps_on_workers=pow_,
                       nvls={"off": False, "on": True}.get(FLAGS.nvls, False if (FLAGS.in_graph or world == 1) else "auto"),
                       optimizer={"kind": FLAGS.optimizer, "lr": FLAGS.learning_rate, "momentum": 0.9})
    eng = PSTrainEngine(MLPSpec(hidden=FLAGS.hidden_units, batch=FLAGS.batch_size), cfg, fabric)
    eng.init_params()
    print("rank %d: ps shards %s, workers %s, placement %s" % (rank, eng.ps_ranks, eng.worker_ranks,
                                                               {k: v.shard for k, v in eng.layout.items()}), flush=True)
    xs, ys = synthetic_mnist(FLAGS.num_train, seed=1)
    local_workers = [r for r in eng.ranks if r in eng.worker_ranks]
    host_feed = FLAGS.input == "host" and len(local_workers) <= 1
    hx = hy = None
    if host_feed and local_workers:
        # shuffled epochs in pinned host memory, the next one gathered in the background (utils/input_pipeline.py); every
        # process builds the same epochs (same seed) and worker w takes batches w, w + W, ... of each
        from distributed_tensorflow_b200.utils.input_pipeline import EpochBatcher
        batcher = EpochBatcher(xs, ys, FLAGS.batch_size, shuffle=True, seed=1)
        hx, hy = batcher.next_epoch()
        woff = eng.worker_ranks.index(local_workers[0])
        per_epoch = batcher.num_batches // cfg.num_workers           # steps this worker takes per epoch
        in_epoch = 0
    elif not host_feed:
        for r in local_workers:
            eng.attach_dataset(r, xs, ys)
    t0 = time.time()
    steps = FLAGS.train_steps
    done = 0
    while done < steps:
        k = min(200, steps - done)
        if host_feed:
            # worker w trains on batches w, w + W, w + 2W, ... like the reference's per-worker next_batch streams; a ps-only
            # process just enqueues its applies (train_loop without batches)
            if local_workers:
                k = min(k, per_epoch - in_epoch)
                losses = eng.train_loop(hx, hy, k, first=in_epoch * cfg.num_workers + woff, stride=cfg.num_workers, depth=4,
                                        prefetch_next=in_epoch + k < per_epoch)
                in_epoch += k
                if in_epoch == per_epoch:
                    hx, hy = batcher.next_epoch()
                    in_epoch = 0
            else:
                eng.train_loop(None, None, k)
        else:
            eng.enqueue_local_steps(k, "dataset")
        done += k
        eng.synchronize()
        loss = float(losses[-1]) if (host_feed and local_workers) else eng.read_loss()
        if loss is not None:
            print("time: %.2fs | rank: %d | local step: %d | loss: %f" % (time.time() - t0, rank, done, loss), flush=True)
    eng.check_errors()
    if not fabric.single_process:
        dist.barrier()
    sd = eng.state_dict()
    if 0 in eng.ranks:
        gs = int(sd["global_step"])
        print("global_step %d after %.2fs  (%.0f samples/s)" % (gs, time.time() - t0,
              cfg.num_workers * FLAGS.batch_size * steps / (time.time() - t0)))
        if not cfg.sync:
            print("staleness:", eng.staleness())
    # name-keyed checkpoint, compatible with distributed_mnist_predict.py (every ps process adds its shard's variables)
    if any(r in eng.ps_ranks for r in eng.ranks):
        if fabric.single_process:
            allsd = sd
        else:
            gathered = [None] * world
            dist.all_gather_object(gathered, sd)
            allsd = {}
            for g in gathered:
                allsd.update(g or {})
        if rank == 0:
            prefix = os.path.join(FLAGS.train_dir, "model.ckpt-%d" % int(allsd["global_step"]))
            write_bundle(prefix, allsd, {"engine": "fabric"})
            update_checkpoint_state(FLAGS.train_dir, prefix)
            print("checkpoint:", prefix)
    elif not fabric.single_process:
        dist.all_gather_object([None] * world, {})
    eng.close()
    if not fabric.single_process:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
