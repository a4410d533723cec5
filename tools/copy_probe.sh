#!/bin/bash
# Host <-> device copy ceiling with N GPUs copying at once (tools/copy_probe.cu): for every allocation kind and
# direction, N processes start at the same wall-clock second and copy for a few seconds.
#   tools/copy_probe.sh <n_gpus> <out.jsonl> [seconds]
set -u
N=${1:-1}; OUT=${2:-gpurun_out/copy_probe.jsonl}; SECS=${3:-3}
HERE=$(cd "$(dirname "$0")" && pwd)
EXE=$HERE/copy_probe
if [ ! -x "$EXE" ]; then nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o "$EXE" "$HERE/copy_probe.cu" || exit 1; fi
: > "$OUT"
echo "{\"nproc\": $(nproc), \"hugepages_total\": $(grep -m1 HugePages_Total /proc/meminfo | awk '{print $2}'), \"thp\": \"$(cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null)\", \"numa_nodes\": $(ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l)}" >> "$OUT"
run() {  # alloc chunk dirs ngpus
  local start=$(( $(date +%s) + 4 + $4 / 2 ))
  for g in $(seq 0 $(( $4 - 1 ))); do "$EXE" $g $start $SECS $1 $2 $3 $4 >> "$OUT" 2>> "$OUT.err" & done
  wait
}
for n in 1 $N; do
  for alloc in pinned wc reg huge; do run $alloc 32 both $n; done
  run pinned 32 h2d $n
  run pinned 32 d2h $n
  run pinned 256 both $n
  [ "$n" = "$N" ] && break
done
cat "$OUT"
