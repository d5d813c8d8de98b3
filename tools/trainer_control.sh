#!/bin/bash
# Side-by-side control on the GPU box: the reference's own trainer with its own CUDA library (oracle/_ref/nts_ref) and the same
# trainer on libnts_b200 (oracle/_ref/nts_b200) on the same cfgs. A cfg that fails in both is a reference problem, not ours.
# usage: bash tools/trainer_control.sh CASE [CASE ...]     (cases = cfg names without the cfg_ prefix / .cfg suffix)
cd "$(dirname "$0")/../oracle/_ref" || exit 1
for c in "$@"; do
  for bin in nts_ref nts_b200; do
    rm -f data/*pre_sample*.bin
    out=$(timeout 300 ./$bin cfg_$c.cfg 2>&1); rc=$?
    accs=$(echo "$out" | grep -o "Train Acc: [0-9.]*" | awk '{print $3}' | tr '\n' ' ')
    why=$(echo "$out" | grep -i "assert\|error\|abort\|fault" | grep -v "AccumulateGrad" | tail -2 | cut -c1-220 | tr '\n' '|')
    echo "$c $bin rc=$rc train_acc=[ $accs] $why"
  done
done
