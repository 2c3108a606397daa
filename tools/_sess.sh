bash tools/gpu_tests.sh r3g "" 0
bash tools/gpu_evidence.sh r3g
