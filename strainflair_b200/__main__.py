"""Console-script targets with the reference's names (setup.cfg:25-31 there): json2csv, compute_strains_abundance."""
import sys

from .cli import compute_strains_abundance_main, json2csv_main  # noqa: F401

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("json2csv", "compute_strains_abundance"):
        cmd = sys.argv.pop(1)
        (json2csv_main if cmd == "json2csv" else compute_strains_abundance_main)()
    else:
        print("usage: python -m strainflair_b200 {json2csv|compute_strains_abundance} [options]")
        sys.exit(2)
