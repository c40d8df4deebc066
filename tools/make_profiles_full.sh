#!/bin/bash
# Only the `ncu --set full` captures of tools/make_profiles.sh (same commands), digested on the box.
set -u
TAG=${1:-r2}
OUT=gpurun_out/prof_${TAG}
mkdir -p $OUT
NCU="ncu --clock-control none"
python tools/perf_conv.py 64 64 8,13,192,257 > $OUT/perf_conv_one.log 2>&1 || exit 1
$NCU --set full --import-source on -k regex:conv3d_umma -s 3 -c 1 -o $OUT/conv64 -f \
    python tools/perf_conv.py 64 64 8,13,192,257 > $OUT/ncu_conv64.log 2>&1
$NCU --set full --import-source on -k regex:conv3d_umma -s 3 -c 1 -o $OUT/conv64_bnstats -f \
    python tools/perf_conv.py 64 64 8,13,192,257 stats > $OUT/ncu_conv64s.log 2>&1
$NCU --set full -k regex:conv3d_tail -s 3 -c 1 -o $OUT/conv_tail -f \
    python tools/perf_conv.py 64 3 8,13,192,257 > $OUT/ncu_tail.log 2>&1
$NCU --set full -k regex:conv3d_umma -s 3 -c 1 -o $OUT/conv_head -f \
    python tools/perf_conv.py 3 64 8,13,192,257 > $OUT/ncu_head.log 2>&1
HPVG_PRECISION=tf32 $NCU --set full -k regex:conv3d_umma -s 6 -c 1 -o $OUT/conv_tf32 -f \
    python tools/perf_conv.py 64 64 8,13,192,257 > $OUT/ncu_conv_tf32.log 2>&1
$NCU --set full -k regex:conv3d_wgrad_kernel -s 40 -c 1 -o $OUT/wgrad -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_wgrad.log 2>&1
$NCU --set full -k regex:bn_train_apply_cl_kernel -s 50 -c 1 -o $OUT/bn_train_apply -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_bnapply.log 2>&1
$NCU --set full -k regex:bn_bwd_apply_cl_kernel -s 3 -c 1 -o $OUT/bn_bwd_apply -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_bnbwd.log 2>&1
$NCU --set full -k regex:lrelu_bwd_cl_kernel -s 10 -c 1 -o $OUT/lrelu_bwd -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_lrelubwd.log 2>&1
HPVG_PROF_DST=$OUT/summary_full python tools/summarize_profiles.py $TAG > $OUT/summarize_full.log 2>&1
find $OUT -name "*.ncu-rep" -delete
ls -la $OUT/summary_full
