#!/bin/bash
# Times the trunk kernel for every variant library under alphaquoridorgnn_b200/variants/ (run on the GPU box).
for so in alphaquoridorgnn_b200/variants/libaqgnn_*.so; do
  n=$(basename $so .so)
  AQ_LIB_PATH=$PWD/$so timeout 300 python bench.py --skip-extra --skip-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n', round(d['ms_per_step'],4), round(d['kernels']['gcn_forward_kernel']['ms'],4), round(d['kernels']['heads_forward_kernel']['ms'],4), round(d['kernels']['legal_mask_kernel']['ms'],4))" || echo "$n FAILED"
done
