set -u
(timeout 120 benchmarks/micro/bin/lb_A R16B2 2; timeout 120 benchmarks/micro/bin/lb_R20 R20B2 2; timeout 120 benchmarks/micro/bin/lb_R24 R24B2 2; timeout 120 benchmarks/micro/bin/lb_R32B1 R32B1 1) 2>&1 | grep -v "single launches"
