"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of a bench.py run) -> profiles/r02_decoder_dram.json,
the per-lane-frame DRAM traffic of the decoder advance kernel that bench.py scales into `roofline.traffic`.
Usage: python tools/dram_capture_to_json.py <capture.csv> <lanes> <frames> <workload> [out.json]
The decoding launch is the advance-kernel launch with the largest traffic (the others are InitDecoding: one frame-0 closure)."""
import csv, json, sys

src, lanes, frames, workload = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
out = sys.argv[5] if len(sys.argv) > 5 else "profiles/r02_decoder_dram.json"
per = {}
for r in csv.reader(open(src, errors="replace")):
    if len(r) < 15 or not r[0].isdigit() or "dec_advance" not in r[4]:
        continue
    per.setdefault(int(r[0]), dict(kernel=r[4], grid=r[8], block=r[7]))[r[12]] = float(r[14])
best = max(per.values(), key=lambda d: d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
rd, wr = best["dram__bytes_read.sum"], best["dram__bytes_write.sum"]
j = dict(workload=workload, kernel=best["kernel"], grid=best["grid"], block=best["block"], lanes=lanes, frames=frames,
         dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_lane_frame=(rd + wr) / (lanes * frames),
         gpu_time_ms_under_ncu=best.get("gpu__time_duration.sum", 0) / 1e6,
         source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (single pass) on bench.py, " + src)
json.dump(j, open(out, "w"), indent=1)
print(json.dumps(j))
