"""Latency of DynamicsWorldModelWrapper.step (one imagined frame per call over the in-place time cache) at the config-4
model, for a few env batch sizes.  Prints one JSON line per batch size and appends them to gpurun_out/env_step_bench.jsonl.

    python scripts/env_step_bench.py [--steps 48] [--batches 1,16,256]
    D4_GRAPH=0 python scripts/env_step_bench.py          # without the CUDA-graph replay of frames (default: on, third episode per batch size is the replay)

Each step = 1 prompted generate() call = 5 transformer passes + reward head (+ terminal head if the model has one); the timed
region is host-visible wall time per step (CUDA events around the whole loop, synchronised), i.e. it includes the Python and
launch overhead that dominates at batch 1.  `d4_launch_count` gives the kernels launched per step."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from dreamer4_b200 import DynamicsWorldModel, DynamicsWorldModelWrapper, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=48)
    ap.add_argument('--warmup', type=int, default=6)
    ap.add_argument('--batches', default='1,16,256')
    ap.add_argument('--precision', default='tf32x3')
    args = ap.parse_args()
    torch.manual_seed(0)
    model = DynamicsWorldModel(**WORKLOADS['config4']['model'], precision=args.precision).cuda()
    lib = _lib.load()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    out = open(os.path.join(ROOT, 'gpurun_out', 'env_step_bench.jsonl'), 'a')
    graphs = os.environ.get('D4_GRAPH', '1') != '0'      # frame graphs are on by default (engine.h: use_graphs)
    for B in [int(b) for b in args.batches.split(',')]:
        # D4_GRAPH=1: a frame graph is keyed by its cache position, run directly the first time it is seen and captured the
        # second time - so episodes 1 and 2 are preparation and episode 3 (same positions again) is the replay being timed
        for episode in range(3 if graphs else 1):
            env = DynamicsWorldModelWrapper(model, num_generation_steps=4)
            env.reset(batch_size=B, seed=1)
            act = torch.zeros(B, 1, dtype=torch.long, device='cuda')
            for _ in range(args.warmup):
                env.step(act)
            torch.cuda.synchronize()
            l0 = lib.d4_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                obs, reward, terminated, truncated, info = env.step(act)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            line = dict(metric='env step latency', unit='ms/step', value=round(ms, 3), env_batch=B, frames_per_s=round(B * 1000. / ms, 1),
                        steps=args.steps, warmup=args.warmup, launches_per_step=(lib.d4_launch_count() - l0) // args.steps,
                        frames_in_cache_at_end=int(env._latents.shape[1]), precision=args.precision, workload='config4 model, num_generation_steps=4',
                        cuda_graphs=graphs, episode=episode + 1, graph_replays=int(lib.d4_graph_replays(model._ctx)) if model._ctx else 0,
                        graph_stats={k: int(lib.d4_debug_get(model._ctx, k.encode())) for k in ('graph_enabled', 'graph_keys', 'graph_captured', 'graph_capture_refused')})
            print(json.dumps(line), flush=True)
            out.write(json.dumps(line) + '\n')
            out.flush()


if __name__ == '__main__':
    main()
