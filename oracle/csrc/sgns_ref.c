/* placeholder translation unit; the gensim-3.8 SGNS restatement lands here */
int orc_sgns_abi(void) { return 0; }
