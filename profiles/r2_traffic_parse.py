"""profiles/traffic.json from the raw csv of profiles/r2_traffic.sh: DRAM bytes (read + write) per frame pair of the HBM-bound
ops captured at the bench batch (74).  usage: python profiles/r2_traffic_parse.py gpurun_out/r2_traffic_raw.csv 74"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2])
hdr = rows[0]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
ur, uw = rows[1][ri], rows[1][wi]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].split("::")[-1].split("<")[0]
    tot.setdefault(name, []).append(float(r[ri].replace(",", "")) * scale[ur] + float(r[wi].replace(",", "")) * scale[uw])
proj = sum(sum(tot[k]) for k in ("project_prep_kernel", "project_tile_kernel", "project_post_kernel"))
out = {
    "project_nn_corr_L1_things_per_frame_pair": round(proj / B),
    "corr2d_L1_things_per_frame_pair": round(sum(tot["corr2d_fwd_diag_kernel"]) / B),
    "event_voxel_things_per_frame_pair": round(sum(tot["event_voxel_int_kernel"]) / len(tot["event_voxel_int_kernel"])),
    "_source": "profiles/r2_traffic_ncu_raw.csv (ncu --set full --clock-control none, batch %d, profiles/r2_traffic.sh): "
               "dram__bytes_read.sum + dram__bytes_write.sum; project_nn_corr_L1 = the four level-1 calls (C2,C3) = (32,32) x2, (81,34), "
               "(96,64), three kernels each; corr2d_L1 = corr2d_fwd_diag_kernel; event_voxel = event_voxel_int_kernel alone (the "
               "41.5 MB zero fill is a cudaMemsetAsync node ncu does not list; its lines and the final write-back leave L2 after "
               "the kernel); bench.py multiplies by its batch" % B,
}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
