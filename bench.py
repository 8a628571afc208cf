#!/usr/bin/env python
"""bench.py — imagined latent frames/s of the Dreamer-4 imagination hot path on B200.

A step = one DreamTrainer iteration (reference dreamer4/trainers.py:1416-1468) on a batch of synthetic dreams:
`generate(H frames, B dreams)` -> `learn_from_experience` -> backward -> clip -> AdamW on the policy and value heads
(+ one gradient all-reduce when N > 1).  Weights are random-init, inputs are the injected noise tensors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.json config 4: 256x256 patch 32 -> 64 latent tokens x 32, dim 512, depth 8, 4 discrete actions
WORKLOADS = {
    'config4': dict(model=dict(dim=512, dim_latent=32, num_latent_tokens=64, depth=8, time_block_every=4, attn_heads=8, attn_dim_head=64,
                               num_discrete_actions=4, predict_terminals=False), batch=2048, horizon=64),
    'config2': dict(model=dict(dim=256, dim_latent=32, num_latent_tokens=32, depth=4, time_block_every=4, attn_heads=8, attn_dim_head=64,
                               num_discrete_actions=(5, 5), predict_terminals=False), batch=1024, horizon=16),
    'config3': dict(model=dict(dim=512, dim_latent=64, num_latent_tokens=64, depth=6, time_block_every=4, attn_heads=8, attn_dim_head=64,
                               num_discrete_actions=4, predict_terminals=False), batch=8192, horizon=32),
    'config1': dict(model=dict(dim=512, dim_latent=32, num_latent_tokens=64, depth=4, time_block_every=4, attn_heads=8, attn_dim_head=64,
                               num_discrete_actions=4, predict_terminals=False), batch=2, horizon=10),
}
METRIC = 'imagined latent frames/sec'
UNIT = 'frames/s'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p.get('hbm_gbs', 6650.), tensor_burst=p.get('bf16_tflops', 1590.), tensor_sustained=p.get('bf16_tflops_sustained', 1400.),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650., tensor_burst=1590., tensor_sustained=1400., source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200', '-i', str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                    power_w_max=max(power) if power else None)


def make_noise(cfg_model, B, H, pinned):
    N, Dl = cfg_model['num_latent_tokens'], cfg_model['dim_latent']
    nda = cfg_model['num_discrete_actions']
    A = sum(nda) if isinstance(nda, (tuple, list)) else nda
    g = torch.Generator().manual_seed(1234)
    lat = torch.randn(H, B, N, Dl, generator=g)
    au = torch.rand(H, B, A, generator=g)
    if pinned:
        lat, au = lat.pin_memory(), au.pin_memory()
    return dict(latent=lat, action_uniform=au, terminal_uniform=torch.zeros(H, 1))


def run_native(args):
    from dreamer4_b200 import DynamicsWorldModel, _lib
    from dreamer4_b200.dist import allreduce_mean_grads_

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    assert torch.cuda.is_available(), 'the native arm needs a CUDA device (no CPU fallback)'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)

    wl = WORKLOADS[args.workload]
    H = args.horizon or wl['horizon']
    # BASELINE.json states config 4 as "64-step dream batch=2048 sharded 8xB200": the GLOBAL dream batch is fixed and split over the
    # ranks (strong scaling, the default).  --scaling weak keeps the per-GPU batch fixed instead (round 1's measurement); with N > 1
    # the strong run also reports a short weak-scaling measurement under "weak".
    global_B = args.batch or wl['batch']
    if args.scaling == 'strong':
        assert global_B % world == 0, f'global dream batch {global_B} does not split over {world} ranks'
        B = global_B // world
    else:
        B = global_B
    torch.manual_seed(0)
    model = DynamicsWorldModel(**wl['model'], precision=args.precision, time_attn_variant=args.variant).to(dev)
    lib = _lib.load()
    policy_optim = torch.optim.AdamW(model.policy_head_parameters(), lr=3e-4)        # trainers.py:1334
    value_optim = torch.optim.AdamW(model.value_head_parameters(), lr=3e-4)
    head_params = model.policy_head_parameters() + model.value_head_parameters()

    # every rank draws the noise of the WHOLE global batch from the same seed and keeps its own shard: the dreams a rank imagines do
    # not depend on how many ranks there are
    host_noise = make_noise(wl['model'], B * world if args.scaling == 'strong' else B, H, pinned=False)
    if args.scaling == 'strong' and world > 1:
        host_noise = {k: (v[:, rank * B:(rank + 1) * B].contiguous() if k != 'terminal_uniform' else v) for k, v in host_noise.items()}
    host_noise = {k: v.pin_memory() for k, v in host_noise.items()}
    dev_noise = {k: v.to(dev) for k, v in host_noise.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for k, v in host_noise.items() if k != 'terminal_uniform')

    phase_events = []

    def mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def step(e2e, host_noise=host_noise, dev_noise=dev_noise, B=B):
        ev = [mark()]
        noise = {k: v.to(dev, non_blocking=True) for k, v in host_noise.items()} if e2e else dev_noise
        exp = model.generate(H, batch_size=B, return_rewards_per_frame=True, return_agent_actions=True,
                             return_log_probs_and_values=True, noise=noise)           # trainers.py:1422-1428
        ev.append(mark())
        pl, vl = model.learn_from_experience(exp)                                     # trainers.py:1430
        pl.backward()
        vl.backward()
        ev.append(mark())
        if dist is not None:            # one flat all-reduce of the head gradients (DDP-equivalent averaging)
            allreduce_mean_grads_(head_params)
        torch.nn.utils.clip_grad_norm_(model.policy_head_parameters(), 0.5)           # trainers.py:1440
        policy_optim.step(); policy_optim.zero_grad()
        torch.nn.utils.clip_grad_norm_(model.value_head_parameters(), 0.5)
        value_optim.step(); value_optim.zero_grad()
        out = torch.stack((pl.detach(), vl.detach(), exp.episode_return.mean()))
        ev.append(mark())
        phase_events.append(ev)
        return out.cpu() if e2e else out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(n, e2e, profile=False, step=step):
        barrier()
        if profile:
            _lib.check(lib.d4_profile(model._ctx, 1))
        l0 = lib.d4_launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        last = None
        for _ in range(n):
            last = step(e2e)
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1)
        launches = lib.d4_launch_count() - l0
        prof = None
        if profile:
            _lib.check(lib.d4_profile(model._ctx, 0))
            buf = (C.c_double * 12)()
            _lib.check(lib.d4_profile_read(model._ctx, buf))
            prof = [list(buf[i * 3:(i + 1) * 3]) for i in range(4)]
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof, last

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    phase_events.clear()
    ms, launches, _, last = timed(args.steps, e2e=False)
    clocks = sampler.stop() if rank == 0 else None
    phases = dict(generate=sum(e[0].elapsed_time(e[1]) for e in phase_events) / args.steps,
                  learn_fwd_bwd=sum(e[1].elapsed_time(e[2]) for e in phase_events) / args.steps,
                  allreduce_clip_adamw=sum(e[2].elapsed_time(e[3]) for e in phase_events) / args.steps)
    step(True)                                   # warm the pinned-copy path
    ms_e2e, _, _, last_e2e = timed(args.steps, e2e=True)
    # kernel-class breakdown: ONE extra step with a CUDA-event pair around every launch of the library (not part of `value`:
    # creating ~10^5 events costs host time)
    prof, ms_prof = None, None
    if not args.no_profile:
        ms_prof, _, prof, _ = timed(1, e2e=False, profile=True)

    # the other scaling regime, measured briefly in the same run (N > 1, strong default only): every rank takes the whole batch
    weak = None
    if args.scaling == 'strong' and world > 1 and not args.no_weak:
        del dev_noise
        model._release()
        torch.cuda.empty_cache()
        wn = make_noise(wl['model'], global_B, H, pinned=False)
        wdev = {k: v.to(dev) for k, v in wn.items()}
        wstep = lambda e2e: step(False, wn, wdev, global_B)
        wstep(False)
        ms_w, _, _, _ = timed(2, e2e=False, step=wstep)
        weak = dict(scaling='weak', dreams_per_gpu=global_B, global_dreams=world * global_B, steps=2, warmup=1, ms_per_step=ms_w / 2,
                    value=world * global_B * H * 2 / (ms_w / 1e3), unit=UNIT)
        del wdev

    frames = world * B * H * args.steps
    value = frames / (ms / 1e3)
    e2e_value = frames / (ms_e2e / 1e3)
    pk = peaks()
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype='tf32' if args.precision == 'tf32' else 'f32',
                data='synthetic', impl='native',
                config=dict(workload=f'{args.workload}: BASELINE.json configs[{dict(config1=0, config2=1, config3=2, config4=3)[args.workload]}] model, '
                                     f'{world * B} dreams x {H} frames' + (f' sharded over {world} GPUs ({B} per GPU)' if world > 1 else '') +
                                     ', 4 denoise + 1 clean pass per frame, generate + learn_from_experience + AdamW',
                            dreams_per_gpu=B, horizon=H, global_dreams=world * B, precision=args.precision,
                            precision_note={'tf32x3': 'fp32 in / fp32 out; every dense product = 3 TF32 tensor-core MMAs (hi/lo split), fp32 accumulate: fp32-level accuracy',
                                            'fp32': 'exact fp32 FMA on CUDA cores', 'tf32': 'single-pass TF32 operands (reduced precision)',
                                            'f16x3': 'fp32 in / fp32 out; transformer dense products = 3 fp16 tensor-core MMAs (hi/lo split of power-of-two pre-scaled operands), '
                                                     'fp32 accumulate: fp32-level accuracy, sampled actions bit-exact vs the oracle over 64 frames at this width; the update (learn_from_experience) on 3xTF32'}[args.precision],
                            time_attn_variant=args.variant,
                            parallelism=f'dp{world} (dream batch sharded, one flat gradient all-reduce)',
                            l2='inputs larger than L2 (KV cache + activations per pass >> 126 MB)' if B * H >= 4096 else 'small problem: L2 resident'),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=12, ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches), clocks=clocks,
                losses=dict(policy=float(last[0]), value=float(last[1])), peaks=pk['source'], phase_ms_per_step=phases)
    if weak is not None:
        line['weak'] = weak
    if prof is not None:
        names = ['gemm', 'time_attn', 'small_attn', 'other']
        total_ms = ms_prof
        shares = {n: prof[i][0] / total_ms for i, n in enumerate(names)}
        gemm_tf = prof[0][2] / (prof[0][0] / 1e3) / 1e12 if prof[0][0] > 0 else 0.
        attn_gbs = prof[1][2] / (prof[1][0] / 1e3) / 1e9 if prof[1][0] > 0 else 0.
        # GEMM roofline: `achieved` counts ALGORITHMIC FLOPs (2*M*N*K per layer).  The fp32-accurate path executes 3 TF32
        # tensor-core products per algorithmic one, and TF32 runs at half the bf16 rate the measured peak is quoted in, so the
        # tensor pipe's own bound for this arithmetic is peak / 6 (peak / 2 for single-pass tf32): `frac_of_arith_bound`.
        terms = {'tf32x3': 3, 'tf32': 1, 'f16x3': 1.5}.get(args.precision)      # f16x3: 3 fp16 MMAs at the bf16 rate = 1.5 TF32 units
        roof_gemm = dict(bound='tensor', kernel='gemm (all linear layers)', achieved=gemm_tf, peak=pk['tensor_sustained'], unit='TFLOP/s',
                         frac=gemm_tf / pk['tensor_sustained'], traffic=None, launches=int(prof[0][1]), share_of_step=shares['gemm'],
                         peak_is='measured cuBLAS bf16 dense, sustained')
        if terms:
            roof_gemm.update(executed_tensor_tflops=gemm_tf * terms, tf32_peak_est=pk['tensor_sustained'] / 2,
                             frac_of_arith_bound=gemm_tf * terms / (pk['tensor_sustained'] / 2),
                             note=f'{args.precision}: {terms} TF32-MMA unit(s) per algorithmic product; TF32 = half the bf16 rate')
        roof_attn = dict(bound='hbm', kernel='time_attn (K1, KV-cache decode)', achieved=attn_gbs, peak=pk['hbm'], unit='GB/s',
                         frac=attn_gbs / pk['hbm'], traffic=None, launches=int(prof[1][1]), share_of_step=shares['time_attn'],
                         )
        # DRAM traffic of one representative launch of each kernel from the committed `ncu --set full` capture (profiles/ncu_summary.json,
        # scripts/gpu_profile_r2.sh).  The summary records the sha256 of the kernel sources it was captured from: a kernel that has
        # changed since gets traffic = null instead of a stale number.
        summ_path = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
        if os.path.exists(summ_path) and args.workload == 'config4' and B == WORKLOADS['config4']['batch'] and args.precision == 'f16x3':
            import hashlib
            summ = json.load(open(summ_path))
            shas = summ.get('kernel_sources_sha256', {})
            cur = lambda rel: hashlib.sha256(open(os.path.join(ROOT, rel), 'rb').read()).hexdigest()[:16]
            k1 = summ.get('k1')
            if k1 and shas.get('k1') == cur('dreamer4_b200/csrc/attn_bulk.cu'):
                roof_attn.update(traffic=k1['dram_bytes'], traffic_launch=k1['launch'], traffic_algorithmic_bytes=k1['algorithmic_bytes'],
                                 traffic_source='profiles/r2_ncu_k1.csv')
            else:
                roof_attn.update(traffic_note='no ncu capture of the current attn_bulk.cu under profiles/')
            ff = next((x for x in summ.get('gemm', []) if 'feed-forward in' in x['layer']), None)
            if ff and shas.get('gemm') == cur('dreamer4_b200/csrc/gemm_f16.cu'):
                roof_gemm.update(traffic=ff['dram_bytes'], traffic_launch=ff['layer'], tensor_pipe_active_pct_ncu=ff['tensor_pipe_active_pct'],
                                 sm_clock_ghz_ncu=ff['sm_clock_ghz'], traffic_source='profiles/r2_ncu_gemm.csv')
            else:
                roof_gemm.update(traffic_note='no ncu capture of the current gemm_f16.cu under profiles/')
        line['roofline'] = roof_gemm if prof[0][0] >= prof[1][0] else roof_attn
        line['roofline_attn'] = roof_attn
        line['roofline_gemm'] = roof_gemm
        line['kernel_class_share'] = shares
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args, steps=1, warmup=0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_sample_shape(args):
    b, h = (int(x) for x in args.cpu_sample.split('x'))
    return b, h


def cpu_baseline(args, steps, warmup):
    """The oracle (CPU restatement of the reference path, oracle/dreamer4_oracle.py) timed on the host cores on a bounded
    sample of the same workload: generate + learn_from_experience + backward."""
    from dreamer4_b200 import DynamicsWorldModel
    from oracle import dreamer4_oracle as O
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b, h = cpu_sample_shape(args)
    torch.manual_seed(0)
    model = DynamicsWorldModel(**wl['model'])
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ocfg = O.config_from_reference_kwargs(**wl['model'])
    noise_t = make_noise(wl['model'], b, h, pinned=False)

    def one():
        noise = O.InjectedNoise(noise_t['latent'], noise_t['action_uniform'], None)
        exp = O.generate(sd, ocfg, h, b, noise=noise)
        sdg = {k: v.clone().requires_grad_(k.startswith(('policy_head', 'value_head')) or k == 'action_embedder.discrete_action_unembed')
               for k, v in sd.items()}
        pl, vl, _ = O.learn_from_experience(sdg, ocfg, exp)
        (pl + vl).backward()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return dict(value=b * h * steps / dt, unit=UNIT, cores=cores, kind='port',
                sample=f'{b} dreams x {h} frames of the {args.workload} model per step, {steps} step(s), {dt:.1f} s, torch {torch.get_num_threads()} threads fp32',
                seconds=dt)


def run_reference(args):
    """Reference arm: the reference's own algorithm on the host cores (the oracle port; the reference package itself cannot be
    imported here — 17 un-vendored dependencies)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    b, h = cpu_sample_shape(args)
    base = cpu_baseline(args, steps=args.steps, warmup=min(args.warmup, 1))
    ms = base['seconds'] / args.steps * 1e3
    line = dict(metric=METRIC, value=base['value'], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=min(args.warmup, 1), ms_per_step=ms,
                higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=f'{args.workload} model, bounded CPU sample: {b} dreams x {h} frames per step', dreams=b, horizon=h),
                cpu_baseline=dict(value=base['value'], unit=UNIT, cores=base['cores'], kind='port', sample=base['sample']),
                e2e=dict(value=base['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='config4', choices=list(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='dreams per GPU (default: the workload\'s)')
    ap.add_argument('--horizon', type=int, default=0)
    ap.add_argument('--precision', default=os.environ.get('D4_BENCH_PRECISION', 'f16x3'), choices=['fp32', 'tf32', 'tf32x3', 'f16x3'],
                    help='f16x3 (default): 3-term fp16 split of pre-scaled operands on tcgen05 kind::f16, fp32-accurate (held to the tf32x3 parity bars at '
                         'the benchmark width and horizon, tests/test_horizon_parity_gpu.py); tf32x3: 3-term TF32 split; fp32: SIMT FMA; tf32: single pass (reduced precision)')
    ap.add_argument('--variant', type=int, default=1, help='K1 kernel variant (0 ld.global staged, 1 cp.async.bulk ring)')
    ap.add_argument('--cpu-sample', default='16x16', help='CPU baseline sample: dreams x frames')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong (default): --batch / the workload batch is the GLOBAL dream batch, sharded over the ranks; weak: per-GPU batch')
    ap.add_argument('--no-weak', action='store_true', help='skip the extra weak-scaling measurement of a strong N > 1 run')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()
