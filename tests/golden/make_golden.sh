#!/bin/sh
# Regenerates reference_golden.json from the reference's generator (needs /root/reference; build container only).
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -I/root/reference -I/root/reference/test make_golden.cpp -o /tmp/make_golden
/tmp/make_golden > reference_golden.json
echo "wrote $(pwd)/reference_golden.json"
