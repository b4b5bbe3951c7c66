#!/bin/bash
# usage: tools/sass.sh <object> <mangled-name regex>  -> SASS of the first matching function, encoding column stripped
obj=$1; pat=$2
fn=$(cuobjdump -elf "$obj" 2>/dev/null | grep -oE "\.text\.[A-Za-z0-9_]+" | sed 's/^\.text\.//' | sort -u | grep -E "$pat" | head -1)
echo "# $fn" >&2
cuobjdump -sass -fun "$fn" "$obj" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/\/\* 0x[0-9a-f]+ \*\///; s/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /'
