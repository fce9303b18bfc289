#!/bin/bash
# end-of-round refresh with the final library: the other BASELINE.json configurations + the launch list of one eager step
bash profiles/r2_measure1.sh 2>&1 | grep -v Warn
bash profiles/r2_launches.sh 2>&1 | tail -2
