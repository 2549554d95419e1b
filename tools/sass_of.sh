#!/bin/bash
# usage: tools/sass_of.sh <lib.so> <mangled-kernel-name>  -> instruction stream without addresses / encodings (for diffing)
cuobjdump -sass -fun "$2" "$1" | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's#^\s+/\*[0-9a-f]{4}\*/\s+##; s#\s*/\* 0x[0-9a-f]+ \*/##; s#;.*$#;#'
