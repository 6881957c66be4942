for spp in 128 512; do for p in 23 22 21; do echo "spp=$spp pool=2^$p"; CRAY_POOL_LOG2=$p SPP=$spp MODES=fast bash tools/bench_modes.sh; done; done
