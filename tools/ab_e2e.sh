# end-to-end A/B of prebuilt library variants (GPU entropy mode, frames read back): dense and realistic, twice each
set -u
for v in "$@"; do
  cp variants/lib_$v.so hvqm4_b200/libhvqm4_b200.so
  for rep in 1 2; do
    echo "== $v e2e dense:     $(timeout 120 python tools/profile_e2e.py 1024 16 1 0 3 1 2>&1 | grep fps | tail -1 | cut -c1-90)"
  done
  echo "== $v e2e realistic: $(timeout 120 python tools/profile_e2e.py 1024 16 1 1 3 1 2>&1 | grep fps | tail -1 | cut -c1-90)"
  echo "== $v recon dense:   $(timeout 120 python tools/profile_recon.py 1024 3 0 2>&1 | tail -1)"
  echo "== $v parity: $(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'gpu_entropy or pipelined or staggered' 2>&1 | tail -1)"
done
