# Aggregation step of benchmark.sh (the reference does the same inline, benchmark.sh:109-176).  Input = the
# ';'-separated statistics lines of lbmcl (lbmcl.hpp:648-669)
#   device;precision;dim;iterations;every;lws;stride;optimize;total_ms;kernels_ms;MLUPS;kernelsMLUPS
# Per group the run with the smallest and the run with the largest total time are dropped (when there are
# more than two) and the remaining runs are averaged.
#
#   awk -F';' -f aggregate.awk stats.csv                      extended format (default):
#        groups on the first eight fields, header line, columns
#        device;precision;dim;iterations;every;lws;stride;optimize;runs;total_ms;kernels_ms;MLUPS;kernelsMLUPS;GBs;roofline
#        (GBs = kernelsMLUPS x 152|304 B per cell update, roofline = GBs / peak; -v peak=<GB/s>, default 6550.7)
#   awk -F';' -v mode=reference -f aggregate.awk stats.csv    the reference's benchmark.csv, for its consumers:
#        groups on device;precision;dim;lws;stride ($1;$2;$3;$6;$7 -- runs that differ only in iterations,
#        every or optimize are merged, as in benchmark.sh:113), no header, 9 columns
#        device;precision;dim;lws;stride;total_ms;kernels_ms;MLUPS;kernelsMLUPS
#        (one deviation: the reference divides by count-2 even for one or two runs -- a division by zero;
#        here one or two runs are averaged as they are)
BEGIN { if (peak == "") peak = 6550.7 }
{
    if (mode == "reference") key = $1 FS $2 FS $3 FS $6 FS $7
    else key = $1 FS $2 FS $3 FS $4 FS $5 FS $6 FS $7 FS $8
    n[key]++
    i = n[key]
    tot[key, i] = $9; ker[key, i] = $10; ml[key, i] = $11; kml[key, i] = $12
    prec[key] = $2
    if (!(key in order)) { order[key] = ++nkeys; keys[nkeys] = key }
}
END {
    if (mode != "reference")
        print "device;precision;dim;iterations;every;lws;stride;optimize;runs;total_ms;kernels_ms;MLUPS;kernelsMLUPS;GBs;roofline"
    for (k = 1; k <= nkeys; k++) {
        key = keys[k]; c = n[key]
        lo = 1; hi = 1
        for (i = 2; i <= c; i++) { if (tot[key, i] < tot[key, lo]) lo = i; if (tot[key, i] > tot[key, hi]) hi = i }
        st = sk = sm = skm = 0; used = 0
        for (i = 1; i <= c; i++) {
            if (c > 2 && (i == lo || i == hi)) continue
            st += tot[key, i]; sk += ker[key, i]; sm += ml[key, i]; skm += kml[key, i]; used++
        }
        if (mode == "reference") {
            printf "%s;%.6g;%.6g;%.6g;%.6g\n", key, st / used, sk / used, sm / used, skm / used
        } else {
            gbs = (skm / used) * (prec[key] == "double" ? 304 : 152) / 1000.0
            printf "%s;%d;%.6g;%.6g;%.6g;%.6g;%.5g;%.3f\n", key, used, st / used, sk / used, sm / used, skm / used, gbs, gbs / peak
        }
    }
}
