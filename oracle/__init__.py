"""Test infrastructure: CPU restatement of the reference hot path and the wrapper over the
reference library built by oracle/refbuild.  Never imported by the product path."""
