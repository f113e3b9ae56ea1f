'''Read an `ncu --set full` report (here, no GPU needed) and print, per hand-written kernel launch, the
numbers profiles/rNN/SUMMARY.md and bench.py's `roofline.traffic` quote:

    python profiles/ncu_extract.py gpurun_out/kernels_r02.ncu-rep [--json profiles/r02/ncu_traffic.json]

duration, dram read / write bytes, DRAM %, tensor-pipe %, warps active %, registers, shared memory.'''
import csv
import io
import json
import subprocess
import sys

WANT = {
    'gpu__time_duration.sum': 'us',
    'dram__bytes_read.sum': 'dram_rd',
    'dram__bytes_write.sum': 'dram_wr',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
    'sm__inst_executed_pipe_tensor.sum': 'tensor_inst',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_pct',
    'launch__registers_per_thread': 'regs',
    'launch__shared_mem_per_block_dynamic': 'smem_dyn',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'lts__t_bytes.sum': 'l2_bytes',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
}
UNIT_SCALE = {'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6,
              'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    out = []
    for r in data:
        rec = {'kernel': r[col['Kernel Name']][:60]}
        for metric, key in WANT.items():
            if metric in col and r[col[metric]] != '':
                v = float(r[col[metric]].replace(',', ''))
                rec[key] = v * UNIT_SCALE.get(units[col[metric]], 1.0)
        out.append(rec)
    for rec in out:
        print(json.dumps(rec))
    if '--json' in sys.argv:
        path = sys.argv[sys.argv.index('--json') + 1]
        tag = sys.argv[sys.argv.index('--tag') + 1] if '--tag' in sys.argv else ''
        agg = {}
        for rec in out:
            name = rec['kernel']
            key = ('k3f' if 'k3f_' in name else 'k3' if 'k3_cross' in name else 'k4' if 'k4_' in name else
                   'k2' if 'k2_' in name else 'k13' if 'k13_' in name else 'k1' if ('k1_sim' in name or 'k1b_sim' in name) else None)
            if key:
                agg.setdefault(key, []).append(rec.get('dram_rd', 0) + rec.get('dram_wr', 0))
        json.dump({'captured_at': tag, 'source': rep.split('/')[-1],
                   'bytes_per_launch': {k: sum(v) / len(v) for k, v in agg.items()}}, open(path, 'w'), indent=1)


if __name__ == '__main__':
    main()
