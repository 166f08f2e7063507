timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
