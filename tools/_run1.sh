cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -q -m gpu -s 2>&1 | grep "chain guided\|passed\|failed\|rror" | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --config 3 --no-cpu --no-full 2>/dev/null | tail -1 | head -c 300; echo
timeout 600 python bench.py --no-cpu --no-full 2>/dev/null | tail -1 | head -c 300; echo
