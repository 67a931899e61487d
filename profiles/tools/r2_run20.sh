set -x
python -m pytest tests -x -q -m gpu -k "bls_api or aggregate_verify or uncompressed" 2>&1 | tail -4
