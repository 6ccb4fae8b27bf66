# rebuilds the library with the attention trace hook compiled in, prints the timeline, then restores the normal build
cd turbo-whisper-workspace_b200/csrc
touch attention_enc.cu
make NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr -Xptxas -v -DATTN_TRACE $ATTN_EXTRA" > /dev/null 2>&1
(cd ../..; timeout 60 python tools/attn_trace.py | tail -${1:-8} | cut -c1-200)
touch attention_enc.cu; make > /dev/null 2>&1
