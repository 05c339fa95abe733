#!/bin/bash
# usage: sass_ops.sh <object or .so> <function substring> [regex]   -- memory / sync instruction order of one kernel
cuobjdump -sass "$1" | awk -v f="$2" '/Function :/{p=index($0,f)>0} p' | grep -E "${3:-LDG|STG|LDS|STS|BAR|EXIT|UTMA|UBLKCP|UTCHMMA|LDTM|RED|ATOM}" 
