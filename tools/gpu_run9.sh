cd $GRAFT_REPO_ROOT
echo "--- plain 300 400"; timeout 120 python tools/attn_probe.py 300 400 2>&1 | tail -n 3
echo "--- plain 150 662"; timeout 120 python tools/attn_probe.py 150 662 2>&1 | tail -n 3
