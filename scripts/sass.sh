#!/bin/bash
# usage: scripts/sass.sh <object-or-so> <function-name-substring>  -> SASS of the first matching function on stdout
f=$(cuobjdump -sass "$1" | grep "Function :" | grep "$2" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$f" "$1"
