# usage: attn_flag_sweep.sh FLAG v1 v2 ...   -- rebuilds attention_enc.cu with -DFLAG=v and times the kernel
cd turbo-whisper-workspace_b200/csrc
flag=$1; shift
for e in "$@"; do
  touch attention_enc.cu
  make NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr -Xptxas -v -D$flag=$e" > /dev/null 2>&1
  echo "=== $flag=$e $(grep -i spill build/attention_enc.ptxas.log | head -1)"
  (cd ../..; timeout 100 python tools/bench_kernels.py 2>&1 | grep attention)
done
touch attention_enc.cu; make > /dev/null 2>&1
