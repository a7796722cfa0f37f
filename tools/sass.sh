#!/bin/bash
# usage: tools/sass.sh <object basename> <mangled-name substring> <out file>
OBJ=/root/repo/page_segmentation_b200/build/$1.o
cuobjdump -sass "$OBJ" 2>/dev/null | awk -v pat="$2" '/Function : /{f=index($0,pat)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > "$3"
wc -l "$3"
