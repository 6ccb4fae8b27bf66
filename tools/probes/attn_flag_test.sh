# usage: attn_flag_test.sh "<nvcc -D flags>" <pytest -k expr>   -- rebuild attention with the flags, run the selected GPU tests, restore
cd turbo-whisper-workspace_b200/csrc
touch attention_enc.cu
make NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr -Xptxas -v $1" > /dev/null 2>&1
(cd ../..; timeout 200 python -m pytest tests -x -q -m gpu -k "$2" 2>&1 | tail -3)
touch attention_enc.cu; make > /dev/null 2>&1
