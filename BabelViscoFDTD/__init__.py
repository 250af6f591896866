"""Drop-in import surface: the names BabelBrain imports from the BabelViscoFDTD package
(SURVEY.md section 8b), re-exported from babelbrain_b200 (B200 CUDA path behind a C ABI)."""
__version__ = '1.2.4+b200'
