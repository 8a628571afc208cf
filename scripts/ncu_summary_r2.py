#!/usr/bin/env python
"""gpurun_out/r2_prof_*.ncu-rep (scripts/gpu_profile_r2.sh) -> profiles/ncu_summary.json + profiles/r2_ncu_*.csv (run where ncu is installed).

The summary records the sha256 of the kernel sources it was captured from; bench.py only quotes `roofline.traffic` from it while those
files are unchanged."""
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r'^(Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|'
                  r'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__cycles_elapsed.avg|sm__cycles_active.avg|sm__warps_active.avg.pct_of_peak_sustained_active|'
                  r'launch__registers_per_thread|launch__shared_mem_per_block_dynamic|smsp__inst_executed.sum$|smsp__issue_active.avg.pct_of_peak_sustained_active|'
                  r'lts__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|'
                  r'l1tex__data_pipe_tc_wavefronts_mem_shared.sum$|l1tex__throughput.avg.pct_of_peak_sustained_elapsed)')


def sha(paths):
    h = hashlib.sha256()
    for p in paths:
        h.update(open(os.path.join(ROOT, p), 'rb').read())
    return h.hexdigest()[:16]


def raw(name):
    path = os.path.join(ROOT, 'gpurun_out', f'r2_prof_{name}.ncu-rep')
    if not os.path.exists(path):
        return None
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    idx = [i for i, h in enumerate(hdr) if KEEP.match(h) and 'per_second' not in h and '.max' not in h and '.min' not in h]
    with open(os.path.join(ROOT, 'profiles', f'r2_ncu_{name}.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    units = rows[1]
    recs = []
    for r in rows[2:]:
        d = {}
        for i in idx:
            try:
                d[hdr[i]] = float(r[i].replace(',', ''))
            except ValueError:
                d[hdr[i]] = r[i]
            d[hdr[i] + '__unit'] = units[i]
        recs.append(d)
    return recs


def us(d):
    v, u = d['gpu__time_duration.sum'], d['gpu__time_duration.sum__unit']
    return v / 1e3 if u.startswith('n') else (v if u.startswith('u') else v * 1e3)


def nbytes(d, key):
    v, u = d[key], d[key + '__unit']
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def main():
    M = 2048 * 15
    layers = [('qkv+gates projection (N=1552,K=512)', 1552, 512, 1), ('attention out-projection (N=512,K=512,residual)', 512, 512, 1),
              ('feed-forward in (N=2730,K=512,GLU)', 2730, 512, 1), ('feed-forward out (N=512,K=1376,residual)', 512, 1376, 1),
              ('pool query projection (N=256,K=512)', 256, 512, 1), ('pool keys/values (9M rows,N=512,K=512)', 512, 512, 9),
              ('pool out-projection (N=512,K=256,residual)', 512, 256, 1)]
    summ = dict(source='scripts/gpu_profile_r2.sh -> scripts/ncu_summary_r2.py (ncu --set full --clock-control none, one B200, config 4, 2048 dreams, f16x3)',
                kernel_sources_sha256=dict(gemm=sha(['dreamer4_b200/csrc/gemm_f16.cu']), k1=sha(['dreamer4_b200/csrc/attn_bulk.cu'])))
    g = raw('gemm')
    if g:
        summ['gemm'] = []
        for d, (name, N, K, mult) in zip(g, layers):
            t = us(d)
            summ['gemm'].append(dict(kernel=d['Kernel Name'][:40], layer=name, duration_us=t, dram_bytes=nbytes(d, 'dram__bytes_read.sum') + nbytes(d, 'dram__bytes_write.sum'),
                                     algorithmic_tflops=2.0 * M * mult * N * K / (t * 1e-6) / 1e12,
                                     tensor_pipe_active_pct=d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                                     sm_clock_ghz=round(d['sm__cycles_elapsed.avg'] / (t * 1e3), 3) if 'sm__cycles_elapsed.avg' in d else None,
                                     l2_throughput_pct=d.get('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
                                     l1_smem_throughput_pct=d.get('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')))
    k = raw('k1')
    if k:
        d = k[0]
        t_ctx = 40
        alg = M * 8 * 64 * 4 * (2 * t_ctx + 4)
        summ['k1'] = dict(kernel=d['Kernel Name'][:40], launch=f't = {t_ctx} cached frames, M = {M} tokens x 8 heads, no append', duration_us=us(d),
                          dram_bytes=nbytes(d, 'dram__bytes_read.sum') + nbytes(d, 'dram__bytes_write.sum'), algorithmic_bytes=alg,
                          dram_throughput_pct=d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'))
    for name, key, alg in (('space', 'space_attn', 2048 * 8 * 15 * 64 * 4 * 5), ('pool', 'pool_attn', None)):
        r = raw(name)
        if r:
            d = r[0]
            summ[key] = dict(kernel=d['Kernel Name'][:40], duration_us=us(d), dram_bytes=nbytes(d, 'dram__bytes_read.sum') + nbytes(d, 'dram__bytes_write.sum'),
                             algorithmic_bytes=alg, dram_throughput_pct=d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                             registers=d.get('launch__registers_per_thread'), warps_active_pct=d.get('sm__warps_active.avg.pct_of_peak_sustained_active'))
    json.dump(summ, open(os.path.join(ROOT, 'profiles', 'ncu_summary.json'), 'w'), indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == '__main__':
    main()
