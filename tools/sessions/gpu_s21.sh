#!/bin/bash
# session 21: the round's final single-GPU record (smoke, GPU tests, bench as the driver runs it, reference arm, launch list, config-5 sweep)
bash tools/gpu_final1.sh
