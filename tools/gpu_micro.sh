#!/bin/bash
cd tools/micro
for n in base ks4 ks1 noflip noblds ng4; do echo -n "$n: "; timeout 60 ./tm_$n 2368 1 2; done
echo -n "base dbg0: "; timeout 60 ./tm_base 2368 0 2
echo -n "base nbuf1: "; timeout 60 ./tm_base 2368 1 1
echo -n "ks4 nbuf1: "; timeout 60 ./tm_ks4 2368 1 1
echo -n "base W=8192: "; timeout 60 ./tm_base 8192 1 2
