# persistent vs split-grid backward call at the per-rank shapes of C3 on 2 / 4 / 8 GPUs (one GPU, isolated calls)
for M in 16384 8192 4096; do
  for P in 0 1; do
    MCLIP_BWD_PERSIST=$P python tools/one_bwd.py 32768 $M 512 2>&1 | tail -1
  done
done
