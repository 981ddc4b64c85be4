#!/bin/bash
# first-contact script for a GPU box: build check, the GPU parity suite, then the smoke entry
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -40
