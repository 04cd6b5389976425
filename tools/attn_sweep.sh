python tools/attn_probe.py
for r in 4 6; do for pf in 2 3; do ATVS_ATTN_R=$r ATVS_ATTN_PF=$pf python tools/attn_probe.py | head -1; done; done
ATVS_ATTN_CTAS=148 python tools/attn_probe.py | head -1
ATVS_ATTN_CTAS=222 python tools/attn_probe.py | head -1
python tools/attn_probe.py 8 64 64 80
python tools/attn_probe.py 2 128 160 240
python bench.py --no-extras --steps 20 > gpurun_out/u3_bench.json 2> gpurun_out/u3_bench.err
