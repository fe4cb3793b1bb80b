#!/bin/bash
# Does any buffer (pinned batch memory, device slot scratch) grow after the warm-up calls?  Prints the allocation
# trace lines of the timed steps (none expected) and the step-time distribution.
cd "$(dirname "$0")/.."
for t in ${THREADS_LIST:-4 8 16}; do
  VGB_ALLOC_TRACE=1 B200SDF_TRACE=1 python scripts/e2e_sweep.py ${WL:-noto} ${STEPS:-300} $t > /tmp/late_$t.out 2> /tmp/late_$t.err
  echo "threads $t: $(tail -1 /tmp/late_$t.out | sed 's/.*: min/min/')"
  awk '/^--- step 0$/{on=1} on && !/^--- step/' /tmp/late_$t.err | sort | uniq -c | head -20
  awk '/took/{ if ($5+0 > 2.0) print "  slow:", $0 }' /tmp/late_$t.err | head
done
