set -u
for i in 1 2 3; do timeout 120 benchmarks/micro/bin/lb_A A$i; done
timeout 120 benchmarks/micro/bin/lb_A Adyn 4 0
timeout 120 benchmarks/micro/bin/lb_R12 R12
