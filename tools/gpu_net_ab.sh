set -u
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_denoise.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15
for th in 0 10 12 14; do echo "== RTO_NET_TILE_H=$th"; RTO_NET_TILE_H=$th timeout 300 bash tools/gpu_check.sh benchq 2>&1 | tail -2; done
