# potrf variants (run under gpurun): recursion, left-looking, look-ahead with several panel widths
for v in "GPL_LOOKAHEAD=0" "GPL_LA_NB=1024" "GPL_LA_NB=2048" "GPL_LA_NB=3072" "GPL_LA_NB=4096"; do
  echo "== $v"
  env $v timeout 200 python tools/gpu_probe.py potrf 2>&1 | grep '"potrf"' | grep -E '"n": (8192|16384|32768)' | cut -c1-90
done
