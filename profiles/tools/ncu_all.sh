#!/bin/bash
# usage: scratch/ncu_all.sh <tag>   (run under gpurun)
set -u
tag=$1
out=gpurun_out/ncu_$tag
mkdir -p $out
python profiles/tools/ncu_wstep.py > $out/plain.log 2>&1 || { echo "plain run failed"; tail -20 $out/plain.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'pcd_kernel|gemm_tn|ce_|transpose_pad|pre_tc|lstm|decode' --csv --log-file $out/launches.csv python profiles/tools/ncu_wstep.py > $out/launches.log 2>&1
prof() {  # name regex skip
  ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o $out/$1 python profiles/tools/ncu_wstep.py > $out/$1.log 2>&1
  ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1.raw.csv 2>/dev/null
  ncu -i $out/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $out/$1.sass.csv.gz
  ncu -i $out/$1.ncu-rep --page source --print-source cuda --csv 2>/dev/null | gzip > $out/$1.cuda.csv.gz
  rm -f $out/$1.ncu-rep
}
for spec in "$@"; do
  case $spec in
    wgrad4) prof wgrad_c4_s1 'KWgrad2<\(int\)4, \(int\)1,' 0 ;;
    wgrad16) prof wgrad_c16_s1 'KWgrad2<\(int\)16, \(int\)1,' 0 ;;
    wgrad8s2) prof wgrad_c8_s2 'KWgrad2<\(int\)8, \(int\)2,' 0 ;;
    gemm) prof gemm_tn 'gemm_tn_3xtf32' 0 ;;
    bwdA4)  prof bwdA_c4_s1  'KBwdA2<\(int\)4, \(int\)1,' 3 ;;
    bwdA16) prof bwdA_c16_s1 'KBwdA2<\(int\)16, \(int\)1,' 3 ;;
    bwdA8s2) prof bwdA_c8_s2 'KBwdA2<\(int\)8, \(int\)2,' 0 ;;
    bwdA16s2) prof bwdA_c16_s2 'KBwdA2<\(int\)16, \(int\)2,' 0 ;;
    bwdB16) prof bwdB_c16 'KBwdB2<\(int\)16,' 3 ;;
    bwdB4) prof bwdB_c4 'KBwdB2<\(int\)4,' 3 ;;
    prebwd64) prof pre_bwd_64 'KPreBwd<\(int\)64' 0 ;;
    preconv64) prof pre_conv_64 'KPreConv<\(int\)64' 0 ;;
    srcgrad) prof source_grad 'KSourceGrad' 18 ;;
    combine4) prof combine_c4 'KCombine<\(int\)4' 3 ;;
    combine16) prof combine_c16 'KCombine<\(int\)16' 3 ;;
    nstats4) prof node_stats_c4 'KNodeStats<\(int\)4' 0 ;;
    fwdA4) prof fwdA_c4_s1 'KFwdA<\(int\)4, \(int\)1,' 0 ;;
    fwdA16) prof fwdA_c16_s1 'KFwdA<\(int\)16, \(int\)1,' 0 ;;
    fwdB4) prof fwdB_c4 'KFwdB<\(int\)4,' 0 ;;
    stembwd) prof stem_bwd 'KStemBwd' 0 ;;
    v4fwdA4) prof fwdA4_c4_s1 'KFwdA4<\(int\)4, \(int\)1,' 0 ;;
    v4fwdA16) prof fwdA4_c16_s1 'KFwdA4<\(int\)16, \(int\)1,' 3 ;;
    v4fwdA8s2) prof fwdA4_c8_s2 'KFwdA4<\(int\)8, \(int\)2,' 0 ;;
    v4fwdB4) prof fwdB4_c4 'KFwdB4<\(int\)4,' 0 ;;
    v4bwdA4) prof bwdA4_c4_s1 'KBwdA4<\(int\)4, \(int\)1,' 3 ;;
    v4bwdA16) prof bwdA4_c16_s1 'KBwdA4<\(int\)16, \(int\)1,' 3 ;;
    v4bwdA8s2) prof bwdA4_c8_s2 'KBwdA4<\(int\)8, \(int\)2,' 0 ;;
    pretcdx) prof pre_tc_dx 'pre_tc_dx_kernel' 0 ;;
    pretcdw) prof pre_tc_dw 'pre_tc_dw_kernel' 0 ;;
    gemmsmall) prof gemm_small 'KSmallGemm<\(int\)32' 0 ;;
    step) ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/step_launches.csv python profiles/tools/ncu_step.py > $out/step_launches.log 2>&1 ;;
  esac
done
ls -la $out
