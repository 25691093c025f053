"""Stand-in for the missing alipay/VCSL submodule (TEST INFRASTRUCTURE)."""
