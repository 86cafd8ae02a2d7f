"""adcc_b200: B200-native (sm_100a) engine for adcc's ADC(n) matvec + Jacobi-Davidson path."""
__version__ = "0.1.0"
