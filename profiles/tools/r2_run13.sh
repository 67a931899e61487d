set -x
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) 2>&1
bash profiles/tools/r2_sanitize.sh 2>&1 | grep -v "^+" | tail -14
