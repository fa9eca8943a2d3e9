#!/bin/bash
# How much does a background `nvidia-smi -lms` poll cost the device-resident RDF call?
Q=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap
echo "no sampler"; python tools/device_gap_probe.py 512 | tail -2
for ms in 200 500 1000; do
  nvidia-smi --id=0 --query-gpu=$Q --format=csv,noheader,nounits -lms $ms > /dev/null 2>&1 &
  SMI=$!
  sleep 1
  echo "nvidia-smi -lms $ms"; python tools/device_gap_probe.py 512 | tail -2
  kill $SMI
done
Q2=index,clocks.sm,clocks.max.sm,clocks_event_reasons.active
nvidia-smi --id=0 --query-gpu=$Q2 --format=csv,noheader,nounits -lms 200 > /dev/null 2>&1 &
SMI=$!
sleep 1
echo "nvidia-smi -lms 200, fewer fields"; python tools/device_gap_probe.py 512 | tail -2
kill $SMI
