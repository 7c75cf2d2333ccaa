// dn-damapper / dn-daligner / dn-dbdust: command-line stand-ins with the argv contract of the Dazzler tools that
// DENTIST's workflow calls directly (Snakefile:1155-1169 `damapper -C ... R Q`, :1143-1151 `daligner ... A B`,
// `DBdust db`), so that the Snakemake rules keep their command lines and output names (SURVEY §8b "what calls it").
// Options come first, databases last; the LAS files land in the current directory exactly like the originals'.
// Built three times from this file with -DDN_TOOL=0|1|2 (build.sh).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "dentist_b200.h"

#ifndef DN_TOOL
#define DN_TOOL 0
#endif

static const char *kName[] = {"dn-damapper", "dn-daligner", "dn-dbdust"};
static const char *kUsage[] = {
    "[-C] [-T<n>] [-e<identity>] [-k<n>] [-t<n>] [-s<spacing>] [-l<min length>] [-m<track>]... <ref:dam|db> <reads:db|dam>",
    "[-A] [-B] [-T<n>] [-e<identity>] [-k<n>] [-w<n>] [-h<n>] [-t<n>] [-s<spacing>] [-l<min length>] [-m<track>]... <A:db|dam> [<B:db|dam>]",
    "[-w<window>] [-t<threshold>] [-m<min length>] <db|dam>",
};

int main(int argc, char **argv) {
    std::vector<const char *> opts, dbs;
    for (int i = 1; i < argc; i++) (argv[i][0] == '-' && argv[i][1] ? opts : dbs).push_back(argv[i]);
    const size_t want_min = DN_TOOL == 0 ? 2 : 1, want_max = DN_TOOL == 2 ? 1 : 2;
    if (dbs.size() < want_min || dbs.size() > want_max) {
        fprintf(stderr, "Usage: %s %s\n", kName[DN_TOOL], kUsage[DN_TOOL]);
        return 1;
    }
    const char *dev = getenv("DN_DEVICE");
    if (dn_init(dev ? atoi(dev) : 0, nullptr) != 0) {
        fprintf(stderr, "%s: %s\n", kName[DN_TOOL], dn_last_error());
        return 1;
    }
    int rc;
    if (DN_TOOL == 0) rc = dn_damap(dbs[0], dbs[1], opts.data(), (int)opts.size(), ".");
    else if (DN_TOOL == 1) rc = dn_dalign(dbs[0], dbs.size() > 1 ? dbs[1] : nullptr, opts.data(), (int)opts.size(), ".");
    else rc = dn_dbdust(dbs[0], opts.data(), (int)opts.size());
    if (rc != 0) fprintf(stderr, "%s: %s\n", kName[DN_TOOL], dn_last_error());
    dn_shutdown();
    return rc == 0 ? 0 : 1;
}
