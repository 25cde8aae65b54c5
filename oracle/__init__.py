"""CPU oracle for the ldsr EM hot path -- TEST INFRASTRUCTURE, never imported by ldsr_b200/."""
