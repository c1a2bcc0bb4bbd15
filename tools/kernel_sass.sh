#!/bin/bash
# SASS of one trace_kernel instantiation without the encoding column (development aid).
# usage: tools/kernel_sass.sh build/traverse.o ILb0ELj0ELi0ELb0E > out.sass   (template-argument substring of the mangled name)
obj=$1; pat=$2
fn=$(cuobjdump -sass "$obj" | grep "Function :" | grep "$pat" | head -1 | awk '{print $3}')
cuobjdump -sass -fun "$fn" "$obj" | grep -E "^ +/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##'
