#!/bin/bash
# Main loop vs epilogue of the wide GEMM: the same shapes with the full epilogue, with the accumulators dropped
# (MMN_WIDE_EPI_DEBUG=none), with the epilogue computed but not stored (nostore), with only its TMEM reads (ldonly) and
# with only dummy arithmetic of about its size (aluonly).  usage: bash profiles/wide_gemm_epilogue_split.sh
for m in full none nostore ldonly aluonly; do
  echo "== MMN_WIDE_EPI_DEBUG=$m"
  MMN_WIDE_EPI_DEBUG=$m python profiles/wide_gemm_bench.py 2>&1
done
