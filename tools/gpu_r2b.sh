#!/bin/bash
# round-2 evidence: ncu launch list + full capture at 1080p and 4K, then the sanitizer passes
bash tools/gpu_prof.sh r02a 1080p_b64
bash tools/gpu_prof.sh r02a_4k 4k_wide_b16
bash tools/gpu_sanitize.sh r02
