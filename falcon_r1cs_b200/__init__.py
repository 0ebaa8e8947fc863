"""B200-native prover backend for falcon-r1cs's Falcon verification circuit."""
