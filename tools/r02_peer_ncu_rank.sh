#!/bin/bash
# torchrun entry: rank 0 runs under ncu (NVLink byte counters of the exchange kernels only), the others plain.
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --clock-control none --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_aperture_peer.sum \
      -k regex:"embed_fwd_vec|peer_put2d" -s 8 -c 16 --csv --log-file gpurun_out/r02_peer_nvl.csv \
      python bench.py "$@"
else
  exec python bench.py "$@"
fi
