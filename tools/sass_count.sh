#!/bin/bash
# usage: tools/sass_count.sh <object.o> [kernel-regex]  -> SASS instruction count per kernel; with a regex also dumps the SASS to /tmp/sass_<n>.txt
obj=$1; re=${2:-.}
cuobjdump -sass $obj | awk -v re="$re" '
/Function : /{name=$3; n=0}
/^\s+\/\*[0-9a-f]{4}\*\//{cnt[name]++}
END{for(k in cnt) if (k ~ re) printf "%6d %s\n", cnt[k], k}' | sort -k2 | c++filt | cut -c1-150
