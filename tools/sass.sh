#!/bin/bash
# tools/sass.sh <kernel-name-substring> [so] -- SASS of one kernel of libf110_b200.so, one instruction per line
SO=${2:-f110_gymnasium_ros2_jazzy_b200/csrc/libf110_b200.so}
cuobjdump -sass "$SO" | awk -v pat="$1" '
  /Function :/ { on = index($0, pat) > 0 }
  on && /^[ \t]+\/\*[0-9a-f]{4}\*\// { sub(/^[ \t]+\/\*/, ""); sub(/\*\/[ \t]+/, " "); sub(/[ \t]*\/\*.*$/, ""); print }'
