# thin-layer experiments: epilogue switched off piece by piece (AVDN_GEMM_DBG bits: 1 no bulk store, 2 no TMEM
# load, 4 no staging stores, 8 no proxy fence, 16 no epilogue barriers, 32 no bulk-group wait)
for d in 0 1 7 63; do AVDN_GEMM_DBG=$d PROBE_ONLY=L3 python tools/gemm_epilogue_probe.py 2>&1 | grep "L3 \|L13" | sed "s/^/dbg=$d /"; done
