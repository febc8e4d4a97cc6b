# Aggregation step of benchmark.sh (the reference does the same in benchmark.sh:109-176): input = the
# ';'-separated statistics lines of lbmcl (lbmcl.hpp:648-669)
#   device;precision;dim;iterations;every;lws;stride;optimize;total_ms;kernels_ms;MLUPS;kernelsMLUPS
# grouped by the first eight fields; per group the run with the smallest and the run with the largest
# total time are dropped (when there are more than two) and the remaining runs are averaged.
# Usage: awk -F';' -f aggregate.awk stats.csv > benchmark.csv
{
    key = $1 FS $2 FS $3 FS $4 FS $5 FS $6 FS $7 FS $8
    n[key]++
    i = n[key]
    tot[key, i] = $9; ker[key, i] = $10; ml[key, i] = $11; kml[key, i] = $12
    if (!(key in order)) { order[key] = ++nkeys; keys[nkeys] = key }
}
END {
    print "device;precision;dim;iterations;every;lws;stride;optimize;runs;total_ms;kernels_ms;MLUPS;kernelsMLUPS"
    for (k = 1; k <= nkeys; k++) {
        key = keys[k]; c = n[key]
        lo = 1; hi = 1
        for (i = 2; i <= c; i++) { if (tot[key, i] < tot[key, lo]) lo = i; if (tot[key, i] > tot[key, hi]) hi = i }
        st = sk = sm = skm = 0; used = 0
        for (i = 1; i <= c; i++) {
            if (c > 2 && (i == lo || i == hi)) continue
            st += tot[key, i]; sk += ker[key, i]; sm += ml[key, i]; skm += kml[key, i]; used++
        }
        printf "%s;%d;%.6g;%.6g;%.6g;%.6g\n", key, used, st / used, sk / used, sm / used, skm / used
    }
}