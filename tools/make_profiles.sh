#!/bin/bash
# Regenerates the profiles/ summaries of a round from a GPU box.  Run under gpurun:
#   gpurun --timeout 900 -- 'bash tools/make_profiles.sh r1'
# then, back on the CPU box:  bash tools/make_profiles.sh r1 summarize
set -u
R=${1:-r1}
if [ "${2:-}" != "summarize" ]; then
  mkdir -p gpurun_out
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_match_tile$' -s 2 -c 1 -o gpurun_out/prof_k2_$R \
      python tools/prof_run.py match 64 text 3 > gpurun_out/prof.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_batch_$R.csv \
      python bench.py --workload batch --files 64 --steps 1 --warmup 3 --workers 1 > gpurun_out/ncu_batch.log 2>&1
  python bench.py 2>&1 | tail -1 > gpurun_out/bench_$R.json
  python bench.py --workload batch 2>&1 | tail -1 > gpurun_out/bench_batch_$R.json
  exit 0
fi
cp gpurun_out/launches_$R.csv profiles/${R}_launches.csv
cp gpurun_out/bench_$R.json profiles/${R}_bench_n1.json
cp gpurun_out/bench_batch_$R.json profiles/${R}_bench_batch_n1.json
cp gpurun_out/launches_batch_$R.csv profiles/${R}_launches_batch.csv
{
  echo "# $R: launch list of the batched small-file path (\`bench.py --workload batch --files 64 --workers 1\`: 64 x 256 KiB files = one 16 MiB group, lzss,huffman compress then decompress)"
  echo
  echo "Command: \`ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv python bench.py --workload batch --files 64 --steps 1 --warmup 3 --workers 1\` (first 260 launches; every kernel runs once per group, the file index is blockIdx.y)"
  echo
  python tools/summarize_launches.py gpurun_out/launches_batch_$R.csv
} > profiles/${R}_launches_batch.md
{
  echo "# $R: launch list of \`bench.py --steps 2 --warmup 3\` (64 MiB text stream: lzss compress+decompress, then the Huffman layer and the K2-only timing)"
  echo
  echo "Command: \`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline\` (numbers printed by that run are not bench values; per-launch times are cold-cache and serialised, compare shares)"
  echo
  python tools/summarize_launches.py gpurun_out/launches_$R.csv
} > profiles/${R}_launches.md
{
  echo "# $R: ncu --set full of k_match_tile (K2), 64 MiB text, one launch"
  echo
  echo "Command: \`ncu --set full --clock-control none --import-source on -k regex:'k_match_tile\$' -s 2 -c 1 python tools/prof_run.py match 64 text 3\`"
  echo
  echo "| section | metric | value |"
  echo "|---|---|---|"
  ncu -i gpurun_out/prof_k2_$R.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
keep=('GPU Speed Of Light Throughput','Compute Workload Analysis','Memory Workload Analysis','Scheduler Statistics','Warp State Statistics','Instruction Statistics','Launch Statistics','Occupancy','Source Counters')
for r in csv.reader(sys.stdin):
    if len(r)>14 and r[12] and r[11] in keep: print(f'| {r[11]} | {r[12]} | {r[14]} {r[13]} |')
"
  echo
  echo "## Raw counters"
  echo '```'
  ncu -i gpurun_out/prof_k2_$R.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; d=dict(zip(h,rows[2]))
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__registers_per_thread','launch__shared_mem_per_block_dynamic']:
    if k in d: print(k, d[k], u[h.index(k)])
"
  echo '```'
  echo
  echo "## Hot source lines (share of warp instructions / of stall samples, average active lanes)"
  echo '```'
  python tools/ncu_lines.py gpurun_out/prof_k2_$R.ncu-rep 25
  echo '```'
} > profiles/${R}_k2_ncu_full.md
echo "profiles/${R}_* written"
