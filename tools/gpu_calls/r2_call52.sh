# co-resident clusters by cluster size when a CTA takes a whole SM (is a 4-CTA multicast cluster viable on 148 SMs?)
./tools/cluster_occupancy | tee gpurun_out/r2_cluster_occupancy.txt
