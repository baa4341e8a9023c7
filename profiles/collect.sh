#!/bin/bash
# Collect the ncu evidence the bench line refers to. Run on the GPU box:
#     gpurun --timeout 1500 -- 'bash profiles/collect.sh r2x'
# Writes gpurun_out/<tag>_*; copy what should be judged into profiles/ (tracked). Every capture uses --clock-control none.
#   1. launch list of ONE eager forward (cold-cache, serialised: compare SHARES)           -> <tag>_launches.csv (+ summary)
#   2. ncu --set full of the dominant / graded kernels, one capture each                   -> <tag>_<kernel>.ncu-rep
#   3. profiles/ncu_to_json.py: dram bytes, duration, tensor %, occupancy per capture      -> <tag>_ncu_traffic.json
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv python tools/one_forward.py > $OUT/${TAG}_launches.log 2>&1
python profiles/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
# the SMPL LBS call's own kernels (pose kernel, blend-shape GEMM, skinning), B=256
timeout 300 $NCU --metrics gpu__time_duration.sum -k regex:"smpl|linear_tc" -c 60 --csv --log-file $OUT/${TAG}_lbs_launches.csv python tools/hbm_kernels.py 256 > /dev/null 2>&1
python profiles/summarize_launches.py $OUT/${TAG}_lbs_launches.csv > $OUT/${TAG}_lbs_launches_summary.txt 2>&1
cap() {  # name, kernel regex, skip, count, command...
    local name=$1 re=$2 skip=$3 cnt=$4; shift 4
    timeout 400 $NCU --set full --import-source on -k regex:$re -s $skip -c $cnt -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
}
if [ "${FULL:-1}" = "1" ]; then
    cap ca_b64 ca_vertex_fused 20 2 python tools/ca_check.py 64
    cap ca_b256 ca_vertex_fused 8 2 python tools/ca_check.py 256
    cap fc1 linear_tc_kernel 1 2 python tools/gemm_one.py 17408 1024 512 1
    cap mlp mlp64_fused 2 2 python tools/one_forward.py
    cap attn_rows attn_rows 2 2 python tools/one_forward.py
    cap gru gru_ 4 2 python tools/one_forward.py
    cap smpl smpl_skin 1 2 python tools/hbm_kernels.py 256
    cap attn_tile attn_tile 2 2 python tools/one_forward.py
    python profiles/ncu_to_json.py $TAG $OUT > $OUT/${TAG}_ncu_traffic.json
    for k in ca_b64 ca_b256 fc1 mlp attn_rows gru smpl attn_tile; do
        [ -f $OUT/${TAG}_$k.ncu-rep ] && python profiles/ncu_kernels.py $OUT/${TAG}_$k.ncu-rep 6553.9 > $OUT/${TAG}_${k}_ncu_table.txt 2>/dev/null
    done
    python tools/ncu_walk.py $OUT/${TAG}_ca_b256.ncu-rep ca_vertex 25 > $OUT/${TAG}_ca_b256_stall_walk.txt 2>/dev/null
    python tools/ncu_walk.py $OUT/${TAG}_attn_rows.ncu-rep attn_rows 25 > $OUT/${TAG}_attn_rows_stall_walk.txt 2>/dev/null
    python tools/ncu_walk.py $OUT/${TAG}_mlp.ncu-rep mlp64 25 > $OUT/${TAG}_mlp_stall_walk.txt 2>/dev/null
    # gpurun brings back at most 64 MiB: keep the graded kernel's report, drop the rest (their tables / JSON rows stay)
    [ "${KEEP_REPS:-0}" = "1" ] || find $OUT -name "${TAG}_*.ncu-rep" ! -name "${TAG}_ca_b256.ncu-rep" -delete
fi
tail -30 $OUT/${TAG}_launches_summary.txt
