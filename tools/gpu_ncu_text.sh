#!/bin/bash
# ncu --set full captures summarised ON THE BOX (the .ncu-rep files are too large to travel back): offline kernels at the bench
# batch and one streaming hop at 16,384 streams.  Writes gpurun_out/ncu_<tag>_*.txt.  Usage: tools/gpu_ncu_text.sh tag
cd "$(dirname "$0")/.."
TAG=$1
mkdir -p gpurun_out /tmp/ncu
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_front|k_back|k_gtconv|k_gtcn" -s 10 -c 10 -f -o /tmp/ncu/off \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu --no-parity --batch 1024 > /tmp/ncu/off.log 2>&1
echo "ncu offline rc=$?"
python tools/ncu_summary.py /tmp/ncu/off.ncu-rep gpurun_out/ncu_${TAG}_summary.txt > /dev/null
for k in k_front k_back "k_gtconv.\(bool\)0" k_gtconv_dec_tc k_gtcn_tc; do
  n=$(echo $k | tr -cd "a-z_0-9"); python tools/ncu_lines.py /tmp/ncu/off.ncu-rep "$k" 0 40 > gpurun_out/ncu_${TAG}_lines_$n.txt 2>&1
done

bash tools/gpu_ncu_stream.sh st 16384 spec > /dev/null 2>&1
python tools/ncu_summary.py /tmp/ncu/prof_st.ncu-rep gpurun_out/ncu_${TAG}_stream16k_summary.txt > /dev/null
for k in ks_gtconv_enc3 ks_gtcn; do python tools/ncu_lines.py /tmp/ncu/prof_st.ncu-rep "$k" 0 30 > gpurun_out/ncu_${TAG}_stream16k_lines_$k.txt 2>&1; done
ls -la gpurun_out | tail -12
