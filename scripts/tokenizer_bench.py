"""Throughput of VideoTokenizer.tokenize / .decode at the config-4 tokenizer (256 x 256, patch 32, dim 512, 64 latents: 128 tokens
per frame), frames per second with CUDA events around whole calls.  One JSON line per (op, batch), appended to
gpurun_out/tokenizer_bench.jsonl.  Written in round 1 for the first hardware run of the tokenizer path (round 2).

    python scripts/tokenizer_bench.py [--frames 16] [--batches 4,32,128] [--precision tf32x3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dreamer4_b200 import VideoTokenizer, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=16)
    ap.add_argument('--batches', default='4,32,128')
    ap.add_argument('--precision', default='tf32x3')
    ap.add_argument('--repeat', type=int, default=3)
    args = ap.parse_args()
    torch.manual_seed(0)
    tok = VideoTokenizer(dim=512, dim_latent=32, patch_size=32, image_size=256, num_latent_tokens=64, precision=args.precision).cuda()
    lib = _lib.load()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    out = open(os.path.join(ROOT, 'gpurun_out', 'tokenizer_bench.jsonl'), 'a')
    T = args.frames
    for B in [int(b) for b in args.batches.split(',')]:
        video = torch.randn(B, 3, T, 256, 256, device='cuda')
        latents = tok.tokenize(video)                       # warm-up (contexts, packing)
        tok.decode(latents)
        for name, fn in (('tokenize', lambda: tok.tokenize(video)), ('decode', lambda: tok.decode(latents))):
            torch.cuda.synchronize()
            l0 = lib.d4_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.repeat):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.repeat
            line = dict(metric=f'tokenizer {name}', unit='frames/s', value=round(B * T * 1000. / ms, 1), ms_per_call=round(ms, 3), batch=B, frames=T,
                        launches_per_frame=(lib.d4_launch_count() - l0) // (args.repeat * T), precision=args.precision,
                        workload='256x256 patch 32 dim 512 depth 4+4, 128 tokens per frame')
            print(json.dumps(line), flush=True)
            out.write(json.dumps(line) + '\n')
            out.flush()


if __name__ == '__main__':
    main()
