#!/bin/bash
# compute-sanitizer passes over the kernels with hand-rolled mbarrier / TMEM / cluster protocols.
# Run on the GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/sanitizer_<tool>_<set>.log; profiles/collect.sh summarises them into profiles/.
# Each pass is bounded by its own timeout so a hung kernel cannot hold the box.
set -u
OUT=${OUT:-gpurun_out}
mkdir -p "$OUT"
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
TC_TESTS="tests/test_gpu_tc.py -k test_linear_tc_matches_fp64"
ATTN_TESTS="tests/test_gpu_parity.py -k 'test_attention_blocks_vs_oracle or test_gru_mid_vs_oracle or test_coevo_blocks_vs_oracle or test_lifter_vs_oracle'"
SPIN_TESTS="tests/test_spin.py -k test_spin_features_vs_reference_golden"
PER=${PER:-240}

run() {  # tool, set name, extra env, tests
    local tool=$1 name=$2 envs=$3 tests=$4
    local log="$OUT/sanitizer_${tool}_${name}.log"
    echo "== $tool $name ($envs)" | tee "$log.head"
    # --launch-timeout 0 etc. are defaults; --error-exitcode makes a finding visible in the pytest exit status too
    eval "env $envs timeout $PER $CS --tool $tool --error-exitcode 86 --log-file $log.raw python -m pytest $tests -x -q -p no:cacheprovider" > "$log.pytest" 2>&1
    local rc=$?
    {
        cat "$log.head"
        echo "exit code: $rc (0 = clean, 86 = sanitizer findings, 124 = timed out)"
        grep -E "ERROR SUMMARY|RACECHECK SUMMARY|========= (Error|Warning|Invalid|Race|Barrier|Hazard)" "$log.raw" 2>/dev/null | sort | uniq -c | sort -rn | head -40
        tail -3 "$log.pytest"
    } > "$log"
    rm -f "$log.head"
    # keep the raw log only when there were findings (they can be large)
    if [ $rc -eq 0 ]; then rm -f "$log.raw"; else head -c 400000 "$log.raw" > "$log.raw.head"; rm -f "$log.raw"; fi
    cat "$log"
}

# (tool, set) pairs: memcheck everywhere; racecheck / synccheck where the protocol is hand-rolled across warps / CTAs
PASSES=${PASSES:-"memcheck:gemm memcheck:gemm_pair memcheck:attn_gru_ca memcheck:gru_persistent memcheck:spin racecheck:gemm_pair racecheck:gemm_pair_relaxed racecheck:attn_gru_ca synccheck:gemm_pair synccheck:attn_gru_ca synccheck:gru_persistent"}
for p in $PASSES; do
    tool=${p%%:*}; name=${p##*:}
    case $name in
        gemm) run $tool gemm "" "$TC_TESTS" ;;
        gemm_pair) run $tool gemm_pair "PMCE_TC_PAIR=2 PMCE_TC_BN=256" "$TC_TESTS" ;;
        gemm_pair_relaxed) run $tool gemm_pair_relaxed "PMCE_TC_PAIR=2 PMCE_TC_BN=256 PMCE_TC_PAIR_RELAXED=1" "$TC_TESTS" ;;
        attn_gru_ca) run $tool attn_gru_ca "" "$ATTN_TESTS" ;;
        gru_persistent) run $tool gru_persistent "PMCE_GRU_PERSISTENT=1" "tests/test_gpu_parity.py -k test_gru_mid_vs_oracle" ;;
        spin) run $tool spin "" "$SPIN_TESTS" ;;
    esac
done
