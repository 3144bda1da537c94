"""Reads the round-end ncu captures in gpurun_out/ (tools/round_end_capture.sh) and writes the
summaries kept under profiles/: the launch list as is, the raw pages of the --set full captures
and one JSON with the metrics the roofline discussion uses."""
import csv, io, json, os, subprocess, sys

TAG = sys.argv[1] if len(sys.argv) > 1 else 'r2'
SRC, DST = 'gpurun_out', 'profiles'
KEYS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_bytes.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_xu.sum',
        'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic']
out = {}
for name in ('attn_cross', 'attn_self', 'gemm_wi', 'gemm_out'):
  rep = os.path.join(SRC, f'{TAG}_{name}.ncu-rep')
  if not os.path.exists(rep):
    continue
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  with open(os.path.join(DST, f'{TAG}_ncu_full_{name}.raw.csv'), 'w') as f:
    f.write(raw)
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units, vals = rows[0], rows[1], rows[2]
  d = {'kernel': vals[hdr.index('Kernel Name')][:80]}
  for k in KEYS:
    if k in hdr:
      i = hdr.index(k)
      d[k] = f'{vals[i]} {units[i]}'.strip()
  out[name] = d
json.dump(out, open(os.path.join(DST, f'{TAG}_ncu_full_summary.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
# launch list
ll = os.path.join(SRC, f'{TAG}_launches.csv')
if os.path.exists(ll):
  lines = [l for l in open(ll) if not l.startswith('==')]
  open(os.path.join(DST, f'{TAG}_launches_final.csv'), 'w').writelines(lines)
  tot = {}
  for row in csv.DictReader(lines):
    n = row['Kernel Name']
    cls = next((c for c in ('gemm', 'attention_combine', 'attention', 'rmsnorm', 'sampler') if c in n), 'other')
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3, 'ms': 1e3}.get(unit, 1.0)
    t = tot.setdefault(cls, [0, 0.0])
    t[0] += 1
    t[1] += v
  print({k: (v[0], round(v[1], 1)) for k, v in tot.items()}, 'total us', round(sum(v[1] for v in tot.values()), 1))
