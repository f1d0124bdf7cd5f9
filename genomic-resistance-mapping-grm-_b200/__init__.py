"""grm_b200 -- B200-native k-mer matrix builder behind GRM's Python wrapper API.

Hot path only: genomes (.fna contigs / .fastq reads) -> genome x k-mer
presence matrix, Kover HDF5 dataset and Ray Surveyor TSV.  All arithmetic runs in
``libgrmkm.so`` (hand-written CUDA for sm_100a, C ABI in ``include/grmkm.h``); there is no
CPU fallback -- importing :mod:`grm_b200.native` without the built library raises.
"""
__version__ = "0.1.0"
