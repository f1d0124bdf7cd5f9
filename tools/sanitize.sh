# usage: bash tools/sanitize.sh <tag> [tools...]
# compute-sanitizer over the small GPU parity tests (look-back chains of k_pack and k_aggregate_cols, dedupe flushes,
# retries, abundance rounds, partial export into peer buffers / owner merge): memcheck, racecheck, synccheck,
# initcheck.  One log per tool under gpurun_out/; tools/sanitize_summary.py condenses them into
# profiles/sanitizer_<tag>.txt.
T=${1:-r02}
shift
TOOLS=${@:-memcheck racecheck synccheck}
SMALL="tests/test_gpu_parity.py::test_known_answer_k4 tests/test_gpu_parity.py::test_random_small \
tests/test_gpu_parity.py::test_text_quirks tests/test_gpu_parity.py::test_degenerate_inputs \
tests/test_gpu_parity.py::test_multiple_files_and_empty_rows tests/test_gpu_parity.py::test_fastq_reads \
tests/test_gpu_parity.py::test_tsv_and_strings tests/test_gpu_parity.py::test_single_pass_parser_chains \
tests/test_gpu_parity.py::test_partial_merge_emulated_ranks tests/test_gpu_parity.py::test_partial_export_into_peer_buffers \
tests/test_gpu_parity.py::test_ordered_emission_equals_gather_path \
tests/test_gpu_units.py::test_dedupe_flushes_and_few_buckets tests/test_gpu_units.py::test_entry_list_too_small_is_retried \
tests/test_gpu_units.py::test_units_across_word_blocks tests/test_gpu_units.py::test_low_complexity_and_ties \
tests/test_gpu_abundance.py::test_rounds_small tests/test_gpu_abundance.py::test_pooled_count_table \
tests/test_gpu_result.py::test_checksum_sum_rows_gram"
for tool in $TOOLS; do
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  [ "$tool" = initcheck ] && extra="--track-unused-memory no"
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool $extra --print-limit 40 --log-file gpurun_out/${T}_san_${tool}.log \
    python -m pytest $SMALL -x -q -k "${SAN_K:-not 1000}" > gpurun_out/${T}_san_${tool}_pytest.log 2>&1
  echo "$tool exit $?" >> gpurun_out/${T}_san_${tool}_pytest.log
  tail -2 gpurun_out/${T}_san_${tool}_pytest.log
  grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/${T}_san_${tool}.log | tail -2
done
