#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): produces the raw profile artefacts under gpurun_out/prof_$TAG/.
# Every ncu pass is preceded by a plain run of the same command that must exit 0 (B200_PROFILING.md).
set -u
TAG=${1:-r2}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
NCU="ncu --clock-control none"
# ---- plain runs first
python bench.py --steps 2 --warmup 3 --no-extras > $OUT/bench_sample_plain.log 2>&1 || exit 1
python tools/train_only.py 1 1 eager 16 > $OUT/train_plain.log 2>&1 || exit 1
python tools/train_vae_only.py 1 0 eager > $OUT/train_vae_plain.log 2>&1 || exit 1
python tools/bench_hbm.py --json $OUT/hbm.json > $OUT/hbm.log 2>&1 || exit 1
python tools/perf_conv.py > $OUT/perf_conv.log 2>&1 || exit 1
# ---- launch lists (per-launch durations; cold-cache, serialised: compare SHARES)
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/launches_sample.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_sample.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/launches_train.csv \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_train.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/launches_train_vae.csv \
    python tools/train_vae_only.py 1 0 eager > $OUT/ncu_train_vae.log 2>&1
# ---- full captures of the dominant kernels
$NCU --set full --import-source on -k regex:conv3d_umma -s 3 -c 1 -o $OUT/conv64 -f \
    python tools/perf_conv.py 64 64 8,13,192,257 > $OUT/ncu_conv64.log 2>&1
$NCU --set full --import-source on -k regex:conv3d_umma -s 3 -c 1 -o $OUT/conv64_bnstats -f \
    python tools/perf_conv.py 64 64 8,13,192,257 stats > $OUT/ncu_conv64s.log 2>&1
$NCU --set full -k regex:conv3d_wgrad_kernel -s 40 -c 1 -o $OUT/wgrad -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_wgrad.log 2>&1
# train-path kernels at the 16-frame finest scale (the last launches of an iteration are the finest scale's)
$NCU --set full -k regex:bn_train_apply_cl_kernel -s 50 -c 1 -o $OUT/bn_train_apply -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_bnapply.log 2>&1
$NCU --set full -k regex:bn_bwd_apply_cl_kernel -s 3 -c 1 -o $OUT/bn_bwd_apply -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_bnbwd.log 2>&1
$NCU --set full -k regex:lrelu_bwd_cl_kernel -s 10 -c 1 -o $OUT/lrelu_bwd -f \
    python tools/train_only.py 1 1 eager 16 > $OUT/ncu_lrelubwd.log 2>&1
# tf32 twin of the conv
HPVG_PRECISION=tf32 $NCU --set full --import-source on -k regex:conv3d_umma -s 3 -c 1 -o $OUT/conv_tf32 -f \
    python tools/perf_conv.py 64 64 8,13,192,257 > $OUT/ncu_conv_tf32.log 2>&1
$NCU --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
    -k regex:"colwalk|frames_to_clip|bn_train_apply|adam_apply|adam_norm|bn_bwd_apply|bn_bwd_reduce|reduce_kernel" -c 14 \
    -o $OUT/hbm_kernels -f python tools/bench_hbm.py --iters 1 --warm 0 > $OUT/ncu_hbm.log 2>&1
# digest ON THE BOX (gpurun copies back at most 64 MiB): tables into $OUT/summary, then drop the raw reports
HPVG_PROF_DST=$OUT/summary python tools/summarize_profiles.py $TAG > $OUT/summarize.log 2>&1
du -sh $OUT; find $OUT -name "*.ncu-rep" -delete
ls -la $OUT $OUT/summary
