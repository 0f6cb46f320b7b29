#!/bin/bash
# rebuild libd4d.so from anywhere; fails loudly
set -e
cd "$(dirname "$0")/.."
python -m diffuman4d_b200.build
