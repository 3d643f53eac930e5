#!/bin/bash
# tools/collect_profiles.sh — here (no GPU): turn what tools/capture_profiles.sh brought back in gpurun_out/ into the tracked
# summaries under profiles/ (round prefix as $1, default r02_final).
set -u
P=${1:-r02_final}; O=gpurun_out
cp $O/final_bench_c2.json profiles/${P}_bench_c2.json
cp $O/final_bench_reference.json profiles/${P}_bench_reference.json
cat $O/final_bench_ntap.json $O/final_bench_planes.json > profiles/${P}_bench_other_kernels.json
grep -v "^==" $O/final_launches_c2.csv | cut -d, -f1,5,9,13-15 | head -40 > profiles/${P}_launches_c2.csv
for k in lanczos2 light comp ars planes; do
  if [ -f $O/final_${k}_ncu.txt ]; then cp $O/final_${k}_ncu.txt profiles/${P}_${k}_ncu.txt; else python tools/ncu_summary.py $O/final_full_$k.ncu-rep profiles/${P}_${k}_ncu.txt > /dev/null; fi
done
python - <<PY
import csv, json, subprocess
raw = subprocess.run(['ncu', '-i', '$O/final_full_lanczos2.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if len(r) > 10]
d = dict(zip(rows[0], rows[2]))
units = dict(zip(rows[0], rows[1]))
def to_bytes(k):
    v = float(d[k].replace(',', '')); u = units[k].lower()
    return int(v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u])
kname = [r for r in csv.reader(raw.splitlines()) if len(r) > 10][2][rows[0].index('Kernel Name')].split('<')[0].split('(')[0].replace('void ', '').strip().split('::')[-1]
entry = {'dram_bytes_read': to_bytes('dram__bytes_read.sum'), 'dram_bytes_write': to_bytes('dram__bytes_write.sum'),
    'frames_per_launch': 32, 'source': 'profiles/${P}_lanczos2_ncu.txt (ncu --set full, one launch of bench.py\'s batch of 32 frames)'}
try:
    t = json.load(open('profiles/traffic.json'))
except Exception:
    t = {}
t[kname] = entry
json.dump(t, open('profiles/traffic.json', 'w'), indent=1)
PY
# SASS opcode histogram of the default headline instantiation
cuobjdump -sass -fun '_ZN4b20022vcs_lanczos2_v2_kernelILi4ELi291ELi0ELi60ELi256ELb0ELb0EEEvNS_6VcsDevENS_11Lanczos2DevENS_13Lanczos2V2DevENS_8VcsBatchE' gstreamer_b200/libb200dsp.so 2>/dev/null \
  | grep -o "^\s*/\*[0-9a-f]*\*/\s*[A-Z0-9_.]*" | awk '{print $2}' | sed 's/\..*//' | grep -v '^$' | sort | uniq -c | sort -rn > profiles/${P}_lanczos2_sass_histogram.txt
ls -la profiles | grep ${P}
