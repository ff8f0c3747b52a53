# /* Copyright (C) 2001
#  * Housemarque Oy
#  * http://www.housemarque.com
#  *
#  * Distributed under the Boost Software License, Version 1.0. (See
#  * accompanying file LICENSE_1_0.txt or copy at
#  * http://www.boost.org/LICENSE_1_0.txt)
#  */
#
# /* Revised by Paul Mensonides (2002)  */
#
# /* See http://www.boost.org/libs/preprocessor for documentation. */
#
# ifndef BOOST_PREPROCESSOR_HPP
# define BOOST_PREPROCESSOR_HPP
#
# include <libint2/boost/preprocessor/library.hpp>
#
# endif
