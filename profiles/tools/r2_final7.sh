set -x
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --skip-extra --no-cpu 2>/dev/null | cut -c1-120
