"""profiles/r02_ncu_headline.json from an ncu metrics pass over ONE full-size launch of the headline kernel.

On the GPU box (tools/gpu_job_headline_ncu.sh):
    ncu --clock-control none -k regex:k_integrate -s 1 -c 1 --csv --metrics <METRICS> --log-file gpurun_out/headline_metrics.csv \
        python tools/quick_perf.py 12500 10000 double auto 1
here:
    python tools/ncu_headline.py gpurun_out/headline_metrics.csv 12500 10000 gpurun_out/headline_csrc_sha.txt
(the last file holds bench.csrc_hash() of the sources the capture ran, written by the same job)

bench.py reports `roofline.traffic` and the pipe-busy fractions from this file only when its `csrc_sha` equals the hash
of the sources the benchmark runs (bench.csrc_hash) and kernel / shard match; otherwise it prints None and says why."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

METRICS = ('gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,'
           'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,'
           'smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum,'
           'smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,'
           'launch__registers_per_thread,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')


def main():
    if len(sys.argv) == 2 and sys.argv[1] == '--metrics':
        print(METRICS); return
    path, n_p, n_s = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    sha = open(sys.argv[4]).read().strip()          # hash of the sources the capture ran (not of whatever is here now)
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) >= 15 and r[0] != 'ID']
    if not rows:
        sys.exit('no metric rows in ' + path)
    name = rows[0][4]
    m = {r[12]: float(r[14].replace(',', '')) for r in rows if r[4] == name}
    # ncu's demangled name -> the label bench.py derives from srb_last_launch
    cfg = re.search(r'Cfg<double, (double|float), (\d), (\d), (\d+), (\d), (\d)>', name)
    fp64, kind, tw, nc = cfg.group(1) == 'double', int(cfg.group(3)), int(cfg.group(4)), int(cfg.group(6))
    class I: pass
    info = I(); info.kind, info.tile_width, info.n_components = kind, tw, nc
    grid, block = rows[0][8], rows[0][7]
    rec = {
        'csrc_sha': sha, 'kernel': bench.kernel_label(info, fp64, kind), 'ncu_kernel_name': name,
        'warp_specialised': 'k_integrate_ws' in name,
        'particles': n_p, 'track_steps': n_s, 'grid': grid, 'block': block,
        'command': f'ncu --clock-control none -k regex:k_integrate -s 1 -c 1 --metrics ... python tools/quick_perf.py {n_p} {n_s} double auto 1',
        'gpu_time_ms_under_ncu': m['gpu__time_duration.sum'] * 1e-6,
        'dram_bytes_per_launch': m['dram__bytes_read.sum'] + m['dram__bytes_write.sum'],
        'dram_read_bytes': m['dram__bytes_read.sum'], 'dram_write_bytes': m['dram__bytes_write.sum'],
        'l2_bytes': m.get('lts__t_bytes.sum'),
        'fp64_pipe_cycles_active_pct': m.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
        'fp64_inst_pct_of_peak': m.get('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
        'inst_executed': m.get('smsp__inst_executed.sum'), 'inst_fp64_pipe': m.get('smsp__inst_executed_pipe_fp64.sum'),
        'inst_tensor_dmma': m.get('sm__inst_executed_pipe_tensor.sum'),
        'issue_active_pct': m.get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'warps_active_pct': m.get('sm__warps_active.avg.pct_of_peak_sustained_active'),
        'registers_per_thread': m.get('launch__registers_per_thread'),
        'smem_bank_conflicts': m.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
    }
    upd = n_p * (n_s - 1) * 256 * 32 * 32
    if rec['inst_tensor_dmma']:
        rec['dmma_fma_per_update'] = rec['inst_tensor_dmma'] * 256 / upd
    out = os.path.join(ROOT, 'profiles', 'r02_ncu_headline.json')
    json.dump(rec, open(out, 'w'), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == '__main__':
    main()
