#!/bin/bash
# Times the drop-in program end to end (scene build + upload + render + tone map + PNG) for the shipped chess configuration.
set -e
cd /root/repo
T=$(mktemp -d); mkdir -p $T/build $T/models/envoMaps
python - <<PY
import sys
sys.path.insert(0,'.')
import b2pt_loader; b2pt_loader.load()
from b2pt import scenes as S
S.write_sky_png('$T/models/envoMaps/sky.png', 2048, 1024)
open('$T/build/conf.json','w').write(S.chess_conf_text(1920,1080,2048,True))
PY
cd $T/build
export B2PT_ASSET_DIR=/root/repo/assets/models
( time /root/repo/final-project-monte-carlo-path-tracer-with-microfacet-bsdf_b200/RayTracing ) > $T/out.log 2> $T/err.log || true
tail -c 300 $T/out.log | tr '\r' '\n' | tail -3; tail -5 $T/err.log
ls -la output.png
( time /root/repo/final-project-monte-carlo-path-tracer-with-microfacet-bsdf_b200/RayTracing --spp 32 ) > $T/out2.log 2> $T/err2.log || true
tail -c 200 $T/out2.log | tr '\r' '\n' | tail -2; tail -5 $T/err2.log
cp output.png /root/repo/gpurun_out/chess_32spp.png
