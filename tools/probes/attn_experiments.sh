cd turbo-whisper-workspace_b200/csrc
for e in 1 2 4; do
  touch attention_enc.cu
  make NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr -Xptxas -v -DATTN_EXPERIMENT=$e" > /dev/null 2>&1
  echo "=== experiment $e"
  (cd ../..; timeout 60 python tools/attn_trace.py | tail -6 | cut -c1-200)
done
