#!/bin/bash
# usage: sass_mix.sh <object> <mangled-kernel>   -- static SASS instruction mix of one kernel
cuobjdump -sass -fun "$2" "$1" > /tmp/sass_mix.txt
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" /tmp/sass_mix.txt | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-25}
echo "total: $(grep -cE '^\s+/\*[0-9a-f]+\*/\s+' /tmp/sass_mix.txt)"
