"""Turn the raw ncu artefacts of scripts/r2_profiles.sh (gpurun_out/) into the committed summaries under profiles/.

    python scripts/summarize_ncu.py            # run in the build container (ncu reads the .ncu-rep without a GPU)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r2'


def short(name):
    for k in ('pool_tile_kernel', 'tile_build_kernel', 'ygemm_compact_bf16_kernel', 'ygemm_compact_kernel', 'ygemm_kernel',
              'pool_quad_kernel', 'pool_list_kernel', 'qlist_build_kernel', 'taps_table_kernel', 'cover_mark_kernel',
              'rowlist_kernel', 'unit_table_kernel', 'table_build_kernel', 'prep_weight_umma_kernel', 'prep_weight_bf16_kernel'):
        if k in name:
            return k
    return name.split('(')[0][-48:]


def launch_list():
    path = os.path.join(SRC, f'{TAG}_launches_raw.csv')
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    iname, imetric, ival, iunit = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
    stats = collections.OrderedDict()
    with open(os.path.join(DST, f'{TAG}_launches_raw.csv'), 'w') as f:
        f.write('kernel,duration_us\n')
        for r in rows:
            if r is hdr or r[imetric] != 'gpu__time_duration.sum':
                continue
            us = float(r[ival].replace(',', '')) / (1e3 if r[iunit] == 'ns' else 1.0)
            stats.setdefault(short(r[iname]), []).append(us)
            f.write(f'"{short(r[iname])}",{us:.2f}\n')
    total = sum(sum(v) for v in stats.values())
    with open(os.path.join(DST, f'{TAG}_launches_summary.csv'), 'w') as f:
        f.write('kernel,launches,mean_us,total_us,share_of_device_time\n')
        for k, v in sorted(stats.items(), key=lambda kv: -sum(kv[1])):
            f.write(f'"{k}",{len(v)},{sum(v) / len(v):.2f},{sum(v):.1f},{sum(v) / total:.4f}\n')
    print(open(os.path.join(DST, f'{TAG}_launches_summary.csv')).read())


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_srcunit_tex.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']


def full_capture(rep, out_name, source):
    path = os.path.join(SRC, rep)
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = {'_source': source, 'kernels': {}}
    for r in rows[2:]:
        name = short(r[hdr.index('Kernel Name')])
        if name in out['kernels']:
            continue
        d = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f'{r[i]} {units[i]}'.strip()
        out['kernels'][name] = d
    with open(os.path.join(DST, out_name), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:3000])


def traffic():
    tpath = os.path.join(DST, 'traffic.json')
    tj = json.load(open(tpath))
    for wl in ('MultiviewC', 'MultiviewX', 'Wildtrack'):
        path = os.path.join(SRC, f'{TAG}_traffic_{wl}.csv')
        if not os.path.exists(path):
            continue
        rows = [r for r in csv.reader(open(path)) if len(r) > 5]
        hdr = next(r for r in rows if 'Kernel Name' in r)
        iname, imetric, ival, iunit = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
        acc = collections.defaultdict(lambda: collections.defaultdict(list))
        for r in rows:
            if r is hdr or not r[imetric].startswith('dram__bytes'):
                continue
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[r[iunit]]
            acc[short(r[iname])][r[imetric]].append(float(r[ival].replace(',', '')) * scale)
        for k, m in acc.items():
            per_launch = sum(sum(v) / len(v) for v in m.values())
            tj.setdefault(wl, {})[k] = int(per_launch / 4)          # quick_time.py runs B = 4 frames per launch
    tj['_comment_r2'] = ('pool_tile_kernel / tile_build_kernel / ygemm_compact_kernel (round 2): `ncu --metrics dram__bytes_read.sum,'
                         'dram__bytes_write.sum -k regex:...` passes of scripts/quick_time.py <workload> 4 0 (scripts/r2_profiles.sh), '
                         'per FRAME (B = 4 per launch)')
    json.dump(tj, open(tpath, 'w'), indent=1)
    print(json.dumps({k: v for k, v in tj.items() if not k.startswith('_')}, indent=1))


if __name__ == '__main__':
    launch_list()
    full_capture(f'{TAG}_fwd_kernels.ncu-rep', f'{TAG}_fwd_kernels_ncu_full.json',
                 'ncu --set full --clock-control none --import-source on -k regex:"pool_tile_kernel|tile_build_kernel|'
                 'ygemm_compact_kernel" -s 3 -c 3 python scripts/quick_time.py MultiviewC 4 0 (scripts/r2_profiles.sh; '
                 'gpurun_out/r2_fwd_kernels.ncu-rep read with ncu -i ... --page raw --csv); B = 4 frames per launch')
    traffic()
